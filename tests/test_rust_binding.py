"""The Rust binding under integration/rust/ cannot be compiled here (no cargo / rustc in the image).  What CAN be checked is
checked: the #[repr(C)] structs of sim-b200-sys/src/lib.rs against the structs of include/esim.h (same fields, same order,
matching types), its extern block against the headers' prototypes, the constants, and that the exporter / shim only use
members that exist in the reference's types (the round-1 sources called `.keys()` on a Vec and passed mask_percentage as the
mask effectiveness)."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB_RS = (ROOT / "integration/rust/sim-b200-sys/src/lib.rs").read_text()
ESIM_H = (ROOT / "include/esim.h").read_text()
POPGEN_H = (ROOT / "include/esim_popgen.h").read_text() + (ROOT / "include/esim_popgen_device.h").read_text()

C_TO_RUST = {"double": "f64", "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "uint8_t": "u8", "uint16_t": "u16", "int": "c_int",
             "const uint32_t*": "*const u32", "const uint8_t*": "*const u8", "const uint16_t*": "*const u16",
             "uint8_t*": "*mut u8", "uint16_t*": "*mut u16", "uint32_t*": "*mut u32"}


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def c_structs(header):
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*\1\s*;", strip_comments(header), flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"(.+?)\s*(\**)\s*(\w+(?:\s*,\s*\w+)*)$", decl)
            ctype = (mm.group(1) + mm.group(2)).replace(" *", "*")
            for name in mm.group(3).split(","):
                fields.append((name.strip(), ctype.strip()))
        out[m.group(1)] = fields
    return out


def rust_structs(text):
    out = {}
    for m in re.finditer(r"pub struct (\w+)\s*\{(.*?)\}", strip_comments(text), flags=re.S):
        fields = re.findall(r"pub (\w+)\s*:\s*([^,}]+?)\s*(?:,|$)", m.group(2).strip() + ",")
        out[m.group(1)] = [(n, " ".join(t.split())) for n, t in fields]
    return out


def test_repr_c_structs_match_the_header():
    c, r = c_structs(ESIM_H), rust_structs(LIB_RS)
    for name in ("EsimConfig", "EsimPopulationSoA", "EsimStepStats", "EsimStateView"):
        assert name in r, name
        want = [(n, C_TO_RUST[t]) for n, t in c[name]]
        assert r[name] == want, "%s:\n rust   %s\n header %s" % (name, r[name], want)


def c_prototypes(header):
    out = {}
    for m in re.finditer(r"^\s*(?:const\s+)?[\w\*]+\s*\*?\s+\**(esim_\w+)\s*\(([^;{]*?)\)\s*;", strip_comments(header), flags=re.M):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        out[m.group(1)] = len(args)
    return out


def test_extern_block_names_functions_that_exist_with_the_same_arity():
    protos = c_prototypes(ESIM_H)
    protos.update(c_prototypes(POPGEN_H))
    block = LIB_RS[LIB_RS.index('extern "C" {'):]
    fns = re.findall(r"pub fn (\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", strip_comments(block), flags=re.S)
    assert len(fns) >= 20
    for name, args in fns:
        assert name in protos, "%s is not declared in include/*.h" % name
        n_args = len([a for a in args.split(",") if a.strip()])
        assert n_args == protos[name], "%s: %d arguments in lib.rs, %d in the header" % (name, n_args, protos[name])


def test_constants_match():
    for name, value in re.findall(r"pub const (ESIM_\w+): \w+ = (-?[0-9A-Fa-fx_]+);", LIB_RS):
        m = re.search(r"#define\s+%s\s+(\S+)" % name, ESIM_H)
        assert m, name
        c_val = int(m.group(1).rstrip("u"), 0)
        assert int(value.replace("_", ""), 0) == c_val, name
    assert "ABI version %s" % re.search(r"#define ESIM_ABI_VERSION (\d+)", ESIM_H).group(1) in LIB_RS


def test_sources_use_members_the_reference_has():
    """Each (pattern, reference file, declaration) pair: the binding may use the member because the reference declares it."""
    ref = Path("/root/reference/sim/src")
    export = (ROOT / "integration/rust/export_b200.rs").read_text()
    shim = (ROOT / "integration/rust/simulator_shim.rs").read_text()
    # the round-1 mistakes must not come back
    assert ".keys()" not in export and ".values()" not in export and ".participants()" not in export
    assert "get_participants()" in export
    assert "cfg.mask_effectiveness = d.mask_effectiveness" in shim and "= d.mask_percentage" not in shim
    if not ref.exists():        # the GPU box has no reference tree: the rest needs it
        return
    uses = [("area.buildings.iter()", "models/output_area.rs", "pub buildings: Vec<"),
            ("area.citizens.iter()", "models/output_area.rs", "pub citizens: Vec<Citizen>"),
            ("get_participants()", "models/building.rs", "pub fn get_participants(&self) -> Vec<CitizenID>"),
            ("school.classes()", "models/building.rs", "pub fn classes(&self) -> &Vec<Class>"),
            ("school.offices()", "models/building.rs", "pub fn offices(&self) -> &Vec<Vec<CitizenID>>"),
            ("b.id().building_index()", "models/building.rs", "pub fn building_index(&self) -> usize"),
            ("output_area_code().index()", "models/output_area.rs", "pub fn index(&self) -> usize"),
            ("c.id().global_index()", "models/citizen.rs", "pub fn global_index(&self) -> usize"),
            ("c.household_code", "models/citizen.rs", "pub household_code: BuildingID"),
            ("c.workplace_code", "models/citizen.rs", "pub workplace_code: BuildingID"),
            ("c.uses_public_transport", "models/citizen.rs", "pub uses_public_transport: bool"),
            ("c.is_mask_compliant", "models/citizen.rs", "pub is_mask_compliant: bool"),
            ("c.disease_status", "models/citizen.rs", "pub disease_status: DiseaseStatus"),
            ("area.id().code()", "models/output_area.rs", "pub fn code(&self) -> &String"),
            ("builder.output_areas", "simulator_builder.rs", "pub output_areas: Vec<OutputArea>")]
    for used, path, decl in uses:
        assert used in export, used
        assert decl in (ref / path).read_text(), "%s: `%s` not found" % (path, decl)
    for field in ("exposure_chance", "exposed_time", "infected_time", "max_time_step", "vaccination_rate", "mask_effectiveness"):
        assert "d.%s" % field in shim and "pub %s:" % field in (ref / "disease.rs").read_text(), field
    assert "pub const BUS_CAPACITY" in (ref / "config.rs").read_text()
    # what else the shim takes from the crate: declared there, imported here
    for used, path, decl in (("builder.disease_model", "simulator_builder.rs", "pub disease_model: DiseaseModel"),
                             ("builder.area_code", "simulator_builder.rs", "pub area_code: String"),
                             ("builder.output_area_lookup", "simulator_builder.rs", "pub output_area_lookup: HashMap<String, u32>"),
                             ("get_memory_usage()?", "config.rs", "pub fn get_memory_usage() -> anyhow::Result<String>"),
                             ("DEBUG_ITERATION_PRINT", "config.rs", "pub const DEBUG_ITERATION_PRINT: usize"),
                             ("SimError::Simulation { message }", "error.rs", "Simulation {\n        message: String,"),
                             ("SimError::InitializationError { message }", "error.rs", "InitializationError {\n        message: String,"),
                             ("SimError::MissingCitizen { citizen_id: message }", "error.rs", "MissingCitizen {\n        citizen_id: String,"),
                             ("SimError::OptionRetrievalFailure { message, key: String::new() }", "error.rs",
                              "OptionRetrievalFailure {\n        message: String,\n        key: String,"),
                             ("SimError::Error { context:", "error.rs", "Error {\n        context: String,")):
        assert used in shim, used
        assert decl in (ref / path).read_text(), "%s: `%s` not found" % (path, decl)
    for line in ("use std::collections::HashMap;", "use std::time::Instant;", "use std::os::raw::c_char;",
                 "use crate::config::{get_memory_usage, DEBUG_ITERATION_PRINT};", "use crate::error::SimError;",
                 "use crate::simulator_builder::SimulatorBuilder;"):
        assert line in shim, line
    assert "pub fn from_code(code: i32, message: String) -> SimError" in shim     # the shim brings what error.rs does not have


def test_error_codes_follow_the_simerror_variants():
    """include/esim.h numbers ESIM_ERR_* after SimError's variants; the shim's from_code maps them back."""
    import re
    header = (ROOT / "include/esim.h").read_text()
    codes = dict(re.findall(r"#define (ESIM_ERR_\w+)\s+(-\d+)", header))
    assert [codes[k] for k in ("ESIM_ERR_DEFAULT", "ESIM_ERR_SIMULATION", "ESIM_ERR_INITIALIZATION", "ESIM_ERR_MISSING_CITIZEN",
                               "ESIM_ERR_OPTION_RETRIEVAL")] == ["-1", "-2", "-3", "-4", "-5"]
    shim = (ROOT / "integration/rust/simulator_shim.rs").read_text()
    arms = dict(re.findall(r"(-\d) => SimError::(\w+)", shim))
    assert arms == {"-1": "Default", "-2": "Simulation", "-3": "InitializationError", "-4": "MissingCitizen", "-5": "OptionRetrievalFailure"}
