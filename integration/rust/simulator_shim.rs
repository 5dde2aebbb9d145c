//! Drop-in body for sim/src/simulator.rs: the public API (`From<SimulatorBuilder>`, `step`, `simulate`, the pub fields
//! `visualisation` reads) stays, everything per time step runs in libesim_b200.so.  Not compiled in this repository.
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::os::raw::c_char;
use std::time::Instant;

use anyhow::Context;
use sim_b200_sys as ffi;

use crate::config::{get_memory_usage, DEBUG_ITERATION_PRINT};      // config.rs:34,42-47
use crate::error::SimError;                                         // error.rs:24-52
use crate::export_b200;                                             // integration/rust/export_b200.rs, added to the crate
use crate::simulator_builder::SimulatorBuilder;                     // simulator_builder.rs:58-69

/// To be added to sim/src/error.rs: the ESIM_ERR_* codes of include/esim.h are numbered after the variants of `SimError`
/// (error.rs:24-52); what has no variant of its own (invalid population, no device, CUDA, communication, IO) arrives as
/// `SimError::Error` with the library's message as its context.
impl SimError {
    pub fn from_code(code: i32, message: String) -> SimError {
        match code {
            -1 => SimError::Default { message },
            -2 => SimError::Simulation { message },
            -3 => SimError::InitializationError { message },
            -4 => SimError::MissingCitizen { citizen_id: message },
            -5 => SimError::OptionRetrievalFailure { message, key: String::new() },
            _ => SimError::Error { context: format!("libesim_b200 error {}: {}", code, message) },
        }
    }
}

pub struct Simulator {
    handle: *mut ffi::EsimSim,
    pub area_code: String,                                   // simulator.rs:88
    pub output_area_lookup: HashMap<String, u32>,            // simulator.rs:90
    area_codes: Vec<CString>,                                // OutputAreaID::code per area index, for exposures.json
    max_time_step: u32,
    n_citizens: usize,
}

fn check(sim: *mut ffi::EsimSim, rc: i32) -> anyhow::Result<i32> {
    if rc >= 0 { return Ok(rc); }
    let msg = unsafe { CStr::from_ptr(ffi::esim_last_error(sim)) }.to_string_lossy().into_owned();
    // ESIM_ERR_* are numbered after the SimError variants (sim/src/error.rs:24-52)
    Err(SimError::from_code(rc, msg).into())
}

impl From<SimulatorBuilder> for Simulator {                 // replaces simulator.rs:601-644
    fn from(builder: SimulatorBuilder) -> Self {
        let soa = export_b200::to_soa(&builder).expect("population cannot be exported");   // kept alive for the call below
        let mut cfg: ffi::EsimConfig = unsafe { std::mem::zeroed() };
        unsafe { ffi::esim_default_config(&mut cfg) };
        let d = &builder.disease_model;                      // disease.rs:97-109
        cfg.exposure_chance = d.exposure_chance; cfg.exposed_time = d.exposed_time as u32; cfg.infected_time = d.infected_time as u32;
        cfg.max_time_step = d.max_time_step as u32; cfg.vaccination_rate = d.vaccination_rate as u32;
        cfg.mask_effectiveness = d.mask_effectiveness;      // 0.70 (disease.rs:127) - NOT mask_percentage (0.8, the compliance share,
                                                            // which only the builder uses: output_area.rs:99,169)
        cfg.bus_capacity = crate::config::BUS_CAPACITY as u32;                      // config.rs:37
        // the intervention thresholds are crate-private constants (interventions.rs:50-57, 71-78): esim_default_config holds
        // the same numbers; if they are ever changed in the crate, set cfg.lockdown_threshold / vaccination_threshold /
        // mask_pt_threshold / mask_everywhere_threshold here (negative = Option::None)
        let mut handle = std::ptr::null_mut();
        // ESIM_GPUS=N: ONE handle drives N GPUs of this machine (esim_create_multi); the caller stays single-threaded like the
        // reference (run/src/main.rs:290-306) and every call below works unchanged on the multi-device handle
        let gpus: u32 = std::env::var("ESIM_GPUS").ok().and_then(|v| v.parse().ok()).unwrap_or(1);
        let rc = if gpus > 1 { unsafe { ffi::esim_create_multi(&cfg, gpus, std::ptr::null(), &mut handle) } }
                 else { unsafe { ffi::esim_create(&cfg, &mut handle) } };
        check(std::ptr::null_mut(), rc).expect("no B200: there is no CPU fallback");
        check(handle, unsafe { ffi::esim_import_population(handle, &soa.view()) }).expect("population rejected");
        Simulator { handle, area_code: builder.area_code, output_area_lookup: builder.output_area_lookup,
                    area_codes: soa.area_codes, max_time_step: cfg.max_time_step, n_citizens: soa.home.len() }
    }
}

impl Simulator {
    /// simulator.rs:131-152
    pub fn step(&mut self) -> anyhow::Result<bool> {
        let mut s = ffi::EsimStepStats::default();
        Ok(check(self.handle, unsafe { ffi::esim_step(self.handle, &mut s) })? == 1)
    }

    /// simulator.rs:108-127: the reference's progress lines show the entries of the time steps 1, 51, 101, ... (loop index %
    /// DEBUG_ITERATION_PRINT == 0) while the disease exists; the loop stays on the device for DEBUG_ITERATION_PRINT steps at a
    /// time, so the line of time step 50 k + 1 is printed when that chunk returns
    pub fn simulate(&mut self, output_name: String) -> anyhow::Result<()> {
        let mut start_time = Instant::now();
        let mut done = 0u32;
        while done < self.max_time_step {
            let mut n = 0u32;
            let alive = check(self.handle, unsafe { ffi::esim_run(self.handle, DEBUG_ITERATION_PRINT as u32, &mut n) })? == 1;
            if n == 0 { break; }
            let mut first = ffi::EsimStepStats::default();   // the first entry of the chunk: `done` is a multiple of 50 here
            check(self.handle, unsafe { ffi::esim_read_stats(self.handle, done, 1, &mut first) })?;
            if first.susceptible != 0 || first.exposed != 0 || first.infected != 0 {   // StatisticEntry::disease_exists
                // the derived Debug of the reference's StatisticEntry (statistics.rs:206-215)
                let entry = format!("StatisticEntry {{ time_step: {}, susceptible: {}, exposed: {}, infected: {}, recovered: {}, vaccinated: {} }}",
                                    first.time_step, first.susceptible, first.exposed, first.infected, first.recovered, first.vaccinated);
                println!("Completed {: >3} time steps, in: {: >6} seconds  Statistics: {},   Memory usage: {}",
                         DEBUG_ITERATION_PRINT, format!("{:.2}", start_time.elapsed().as_secs_f64()), entry, get_memory_usage()?);
                start_time = Instant::now();
            }
            done += n;
            if !alive { break; }
        }
        let dir = CString::new(output_name).context("output name")?;
        let codes: Vec<*const c_char> = self.area_codes.iter().map(|c| c.as_ptr()).collect();
        check(self.handle, unsafe { ffi::esim_dump_statistics(self.handle, dir.as_ptr(), codes.as_ptr()) })?;   // statistics.rs:113-150
        Ok(())
    }

    /// What `visualisation` reads from `output_areas[*].citizens` (citizen_connections.rs:37-60): refreshed on demand
    pub fn citizen_states(&mut self) -> anyhow::Result<(Vec<u8>, Vec<u16>, Vec<u32>)> {
        let (mut status, mut timer, mut cur) = (vec![0u8; self.n_citizens], vec![0u16; self.n_citizens], vec![0u32; self.n_citizens]);
        let mut view = ffi::EsimStateView { status: status.as_mut_ptr(), timer: timer.as_mut_ptr(), current_bldg: cur.as_mut_ptr(),
                                            on_pt: std::ptr::null_mut(), vax_eligible: std::ptr::null_mut() };
        check(self.handle, unsafe { ffi::esim_read_state(self.handle, &mut view) })?;
        Ok((status, timer, cur))
    }
}

impl Drop for Simulator { fn drop(&mut self) { unsafe { ffi::esim_destroy(self.handle) } } }
