"""The Rust side of the boundary (rust_shim/) cannot be compiled here (no Rust toolchain in the image), so these checks
keep it from going stale: ffi.rs must be what scripts/gen_rust_ffi.py generates from include/minimcmc.h (every entry
point, no more, no fewer), build.rs must list the Makefile's translation units with the same -fmad split and link the
libraries the Makefile links, every ffi call made by the wrapper modules must exist in the header with the same number
of arguments, and the files must at least be balanced Rust text."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "rust_shim")
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_ffi_rs_is_generated_from_the_header():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_rust_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    import gen_rust_ffi as g
    from test_cabi_symbols import declared_functions

    names = [n for _, n, _ in g.prototypes()]
    assert sorted(names) == sorted(declared_functions()), "generator and symbol test disagree on the header's entry points"
    text = open(os.path.join(SHIM, "src", "ffi.rs")).read()
    for n in names:
        assert re.search(rf"pub fn {n}\(", text), n


def test_build_rs_matches_the_makefile():
    mk = open(os.path.join(ROOT, "mini_mcmc_b200", "csrc", "Makefile")).read()
    rs = open(os.path.join(SHIM, "build.rs")).read()
    srcs = set(re.findall(r"\b(mmc_\w+\.cu)\b", mk.split("OBJS")[0]))
    nofmad = set(m + ".cu" for m in re.findall(r"\$\(OBJDIR\)/(mmc_\w+)\.o", mk.split("-fmad=false")[0].split("all: $(OUT)")[1]))
    listed = dict(re.findall(r'\("(mmc_\w+\.cu)", (true|false)\)', rs))
    assert set(listed) == srcs, (sorted(srcs), sorted(listed))
    assert {s for s, f in listed.items() if f == "false"} == nofmad
    for cpp in re.findall(r"\b(mmc_\w+\.cpp)\b", mk):
        assert f'"{cpp}"' in rs
    for lib in ("cudart", "cuda", "stdc++", "pthread"):
        assert f"dylib={lib}" in rs
    for src in list(listed) + ["mmc_host_widen.cpp", "mmc_sink_csv.cpp"]:
        assert os.path.exists(os.path.join(ROOT, "mini_mcmc_b200", "csrc", src)), src


def _strip(text):
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r'"(?:\\.|[^"\\])*"', '""', text)
    return re.sub(r"'(?:\\.|[^'\\])'", "' '", text)


def test_rust_sources_are_balanced_and_call_existing_entry_points():
    import gen_rust_ffi as g

    arity = {n: len(p) for _, n, p in g.prototypes()}
    mods = [f for f in os.listdir(os.path.join(SHIM, "src")) if f.endswith(".rs")]
    assert {"lib.rs", "ffi.rs", "core.rs", "metropolis_hastings.rs", "hmc.rs", "nuts.rs", "stats.rs", "gibbs.rs", "distributions.rs"} <= set(mods)
    lib = open(os.path.join(SHIM, "src", "lib.rs")).read()
    for f in mods:
        if f != "lib.rs":
            assert f"pub mod {f[:-3]};" in lib, f
    for f in mods + ["../build.rs"]:
        text = _strip(open(os.path.join(SHIM, "src", f)).read())
        for a, b in ("()", "[]", "{}"):
            assert text.count(a) == text.count(b), f"{f}: unbalanced {a}{b}"
        if f in ("ffi.rs", "../build.rs"):
            continue
        for m in re.finditer(r"\b(mmc_[a-z0-9_]+)\s*\(", text):
            name = m.group(1)
            if name not in arity:
                continue   # a type (mmc_run_stats::default()) or constructor, not an entry point
            # count top-level commas of the call's argument list
            depth, i, commas, empty = 0, m.end(), 0, True
            while True:
                ch = text[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    if depth == 0:
                        break
                    depth -= 1
                elif ch == "," and depth == 0:
                    commas += 1
                if not ch.isspace():
                    empty = False
                i += 1
            n_args = 0 if empty else commas + 1
            assert n_args == arity[name], f"{f}: {name} called with {n_args} arguments, the header declares {arity[name]}"


def test_reference_surface_is_covered():
    """Every public entry of the path that SURVEY section 8(a) lists has a b200 body in the shim."""
    want = {
        "metropolis_hastings.rs": ["pub fn new", "pub fn seed", "fn run_device", "fn run_progress_device"],
        "hmc.rs": ["pub fn new", "pub fn set_seed", "pub fn step", "pub fn run", "pub fn run_progress"],
        "nuts.rs": ["pub fn new", "pub fn set_seed", "pub fn run", "pub fn run_progress"],
        "gibbs.rs": ["pub fn new", "pub fn set_seed", "fn run_device"],
        "stats.rs": ["pub fn split_rhat_mean_ess", "pub fn basic_stats", "pub struct RunStats", "split_rhat_mean_ess_sharded"],
        "core.rs": ["pub trait DeviceRunner", "pub fn init_det", "pub fn init_with_seed"],
        "distributions.rs": ["pub trait DeviceTarget", "pub trait DeviceProposal", "RosenbrockND", "DiffableGaussian2D", "PoissonTarget"],
    }
    for f, items in want.items():
        text = open(os.path.join(SHIM, "src", f)).read()
        for it in items:
            assert it in text, f"{f}: {it}"
