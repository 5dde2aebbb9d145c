//! Exporter of a built `SimulatorBuilder` (sim/src/simulator_builder.rs:58-69) to the arrays of `EsimPopulationSoA` and to
//! the binary population file read by `esim_population_load` (format: epidemicsimulator_b200/csrc/population_io.cpp).
//! Not compiled in this repository (no Rust toolchain in the image); tests/test_population_file.py pins the format from the
//! C / Python side, including the checksum and the rejection of damaged files.
use std::collections::HashMap;
use std::ffi::CString;
use std::io::Write;
use std::path::Path;

use sim_b200_sys as ffi;

use crate::disease::DiseaseStatus;                                  // disease.rs:36-44
use crate::models::building::{Building, BuildingID, School, Workplace};   // building.rs:62-67,125-140,220-232,330-342
use crate::models::citizen::{Citizen, CitizenID};                   // citizen.rs:51-67,109-135
use crate::simulator_builder::SimulatorBuilder;                     // simulator_builder.rs:58-69

pub struct Soa {
    pub home: Vec<u32>, pub work: Vec<u32>, pub room: Vec<u32>, pub flags: Vec<u8>, pub status: Vec<u8>, pub timer: Vec<u16>,
    pub bldg_area: Vec<u32>, pub bldg_type: Vec<u8>, pub room_bldg: Vec<u32>, pub area_first: Vec<u32>, pub area_codes: Vec<CString>,
}

impl Soa {
    pub fn view(&self) -> ffi::EsimPopulationSoA {
        ffi::EsimPopulationSoA {
            n_citizens: self.home.len() as u32, n_areas: (self.area_first.len() - 1) as u32, n_buildings: self.bldg_area.len() as u32,
            n_rooms: self.room_bldg.len() as u32, n_global_citizens: 0, n_shared_bldgs: 0, n_shared_rooms: 0, n_shards: 0,
            home_bldg: self.home.as_ptr(), work_bldg: self.work.as_ptr(), room: self.room.as_ptr(), age: std::ptr::null(),
            occupation: std::ptr::null(), flags: self.flags.as_ptr(), status: self.status.as_ptr(), timer: self.timer.as_ptr(),
            global_id: std::ptr::null(), bldg_area: self.bldg_area.as_ptr(), bldg_type: self.bldg_type.as_ptr(), room_bldg: self.room_bldg.as_ptr(),
        }
    }
}

/// Walks `output_areas` in index order (INTEGRATION.md section 6 explains the numbering).
/// Reference types used: `OutputArea { citizens: Vec<Citizen>, buildings: Vec<Box<dyn Building + Sync + Send>>, .. }`
/// (output_area.rs:85-100), `Building::{id, as_any}` (building.rs:125-140), `School::{classes, offices}` (building.rs:444-450),
/// `Class::get_participants` (building.rs:318, returns an owned Vec), `Citizen` pub fields (citizen.rs:109-135).
pub fn to_soa(builder: &SimulatorBuilder) -> anyhow::Result<Soa> {
    let areas = &builder.output_areas;
    let mut bldg_offset = vec![0u32; areas.len() + 1];
    for (a, area) in areas.iter().enumerate() { bldg_offset[a + 1] = bldg_offset[a] + area.buildings.len() as u32; }
    let mut s = Soa { home: vec![], work: vec![], room: vec![], flags: vec![], status: vec![], timer: vec![], bldg_area: vec![],
                      bldg_type: vec![], room_bldg: vec![], area_first: vec![0], area_codes: vec![] };
    let mut room_of: HashMap<CitizenID, u32> = HashMap::new();
    for (a, area) in areas.iter().enumerate() {
        // `buildings` is a Vec indexed by BuildingID::building_index (building.rs:62-67, :107): keep that order
        for (position, b) in area.buildings.iter().enumerate() {
            anyhow::ensure!(b.id().building_index() == position, "building {} of area {} is filed under index {}", b.id().building_index(), a, position);
            let global = bldg_offset[a] + position as u32;
            s.bldg_area.push(a as u32);
            if let Some(school) = b.as_any().downcast_ref::<School>() {
                s.bldg_type.push(2);
                for class in school.classes() {                                    // building.rs:307-342: classes, then offices
                    let r = s.room_bldg.len() as u32; s.room_bldg.push(global);
                    for c in class.get_participants() { room_of.insert(c, r); }    // owned Vec<CitizenID> (building.rs:318)
                }
                for office in school.offices() {
                    let r = s.room_bldg.len() as u32; s.room_bldg.push(global);
                    for c in office.iter() { room_of.insert(c.clone(), r); }
                }
            } else if b.as_any().downcast_ref::<Workplace>().is_some() { s.bldg_type.push(1) } else { s.bldg_type.push(0) }
        }
    }
    let cell = |id: &BuildingID| bldg_offset[id.output_area_code().index()] + id.building_index() as u32;
    // citizen i of the arrays must be CitizenID::global_index() == i (citizen.rs:51-67): that index keys the random stream
    // and is what esim_read_state returns positions for.  The builder numbers citizens area by area, so walking the areas in
    // index order and the citizens of an area by global index gives 0, 1, 2, ...; anything else is refused.
    for area in areas.iter() {
        let mut cs: Vec<&Citizen> = area.citizens.iter().collect();
        cs.sort_by_key(|c| c.id().global_index());
        for c in cs {
            anyhow::ensure!(c.id().global_index() == s.home.len(), "citizens are not numbered area by area");
            s.home.push(cell(&c.household_code));
            s.work.push(cell(&c.workplace_code));
            s.room.push(*room_of.get(&c.id()).unwrap_or(&ffi::ESIM_NO_ROOM));
            s.flags.push(c.uses_public_transport as u8 | (c.is_mask_compliant as u8) << 1);   // ESIM_FLAG_USES_PT | ESIM_FLAG_MASK_COMPLIANT
            let (st, t) = match c.disease_status {                                 // disease.rs:36-44
                DiseaseStatus::Susceptible => (0u8, 0u16), DiseaseStatus::Exposed(t) => (1, t), DiseaseStatus::Infected(t) => (2, t),
                DiseaseStatus::Recovered => (3, 0), DiseaseStatus::Vaccinated => (4, 0) };
            s.status.push(st); s.timer.push(t);
        }
        s.area_first.push(s.home.len() as u32);
        s.area_codes.push(CString::new(area.id().code().as_str()).unwrap());       // output_area.rs:42-54
    }
    Ok(s)
}

// ---- the file: 128-byte header, arrays padded to 64 bytes, FNV-1a 64 trailer (population_io.cpp) ------------------------------
const HAS_STATUS: u32 = 4; const HAS_TIMER: u32 = 8; const HAS_AREA_OFFSETS: u32 = 32; const HAS_AREA_CODES: u32 = 64;

struct Hashing<W: Write> { w: W, hash: u64, written: u64 }
impl<W: Write> Hashing<W> {
    fn raw(&mut self, bytes: &[u8]) -> std::io::Result<()> {
        for b in bytes { self.hash ^= *b as u64; self.hash = self.hash.wrapping_mul(1099511628211); }
        self.written += bytes.len() as u64;
        self.w.write_all(bytes)
    }
    fn array<T: Copy>(&mut self, data: &[T]) -> std::io::Result<()> {   // little-endian host assumed, like the reader
        let bytes = unsafe { std::slice::from_raw_parts(data.as_ptr() as *const u8, std::mem::size_of_val(data)) };
        self.raw(bytes)?;
        let pad = (64 - bytes.len() % 64) % 64;
        self.raw(&[0u8; 64][..pad])
    }
}
fn pad64(n: usize) -> u64 { ((n + 63) & !63) as u64 }

pub fn write_esimpop(path: &Path, s: &Soa) -> anyhow::Result<()> {
    let (n, a, b, r) = (s.home.len(), s.area_first.len() - 1, s.bldg_area.len(), s.room_bldg.len());
    let mut code_off = vec![0u32]; let mut code_blob: Vec<u8> = vec![];
    for c in &s.area_codes { code_blob.extend_from_slice(c.as_bytes()); code_off.push(code_blob.len() as u32); }
    let payload = 3 * pad64(n * 4) + pad64(n) /*flags*/ + pad64(n) /*status*/ + pad64(n * 2) /*timer*/
        + pad64(b * 4) + pad64(b) + pad64(r * 4) + pad64((a + 1) * 4) + pad64((a + 1) * 4) + pad64(code_blob.len());
    let mut header = Vec::with_capacity(128);
    header.extend_from_slice(b"ESIMPOP\x01");
    for v in [1u32, 128, n as u32, a as u32, b as u32, r as u32, 0, 0, 0, 0,
              HAS_STATUS | HAS_TIMER | HAS_AREA_OFFSETS | HAS_AREA_CODES, code_blob.len() as u32] { header.extend_from_slice(&v.to_le_bytes()); }
    header.extend_from_slice(&payload.to_le_bytes());
    header.resize(128, 0);
    let mut f = Hashing { w: std::io::BufWriter::new(std::fs::File::create(path)?), hash: 14695981039346656037, written: 0 };
    f.raw(&header)?;
    f.array(&s.home)?; f.array(&s.work)?; f.array(&s.room)?; f.array(&s.flags)?; f.array(&s.status)?; f.array(&s.timer)?;
    f.array(&s.bldg_area)?; f.array(&s.bldg_type)?; f.array(&s.room_bldg)?; f.array(&s.area_first)?;
    f.array(&code_off)?; f.array(&code_blob)?;
    anyhow::ensure!(f.written == 128 + payload, "size mismatch");
    let digest = f.hash;
    f.w.write_all(&digest.to_le_bytes())?;
    Ok(())
}
