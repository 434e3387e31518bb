"""oracle/build_ref.py -- byte-compile the reference's own modules of the hot path into oracle/_ref/.

The reference is pure Python: its "build" is CPython's compiler.  This recipe compiles the modules where
they lie under /root/reference (read-only; nothing is copied) and writes ONLY the resulting .pyc files
into oracle/_ref/ (git-ignored, not gpurun-ignored: like our own built .so files they travel to the GPU
box, where /root/reference does not exist; they are named <module>.pycode because the snapshot drops
*.pyc).  Loading those code objects gives the UNMODIFIED reference implementation on the GPU box's host
cores:

    bench.py --impl reference      times it (cpu_baseline.kind = "reference")
    tests/test_oracle_cpu.py       re-checks oracle/fwm_oracle.py against it, bit for bit, wherever
                                   oracle/_ref exists (so the pin is re-verified on the GPU box too)

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load oracle/_ref.
`load()` returns the imported modules (matplotlib, which scan_mismtach imports for its plots and which is
not installed, is replaced by an inert stub first).
"""
from __future__ import annotations

import py_compile
import sys
import types
from pathlib import Path

REF_SRC = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "_ref"
MODULES = ["constants", "config", "frequency_plan", "dispersion", "phase_matching", "parameters",
           "integrators", "yaman_model", "simulation", "plotting", "scan_mismtach", "io_fwm"]


def build(force: bool = False) -> bool:
    """Compile the reference modules into oracle/_ref/*.pyc.  Returns True when oracle/_ref is usable."""
    if not REF_SRC.exists():
        return available()
    OUT.mkdir(exist_ok=True)
    for name in MODULES:
        src, dst = REF_SRC / f"{name}.py", OUT / f"{name}.pycode"
        if force or not dst.exists() or dst.stat().st_mtime < src.stat().st_mtime:
            # unchecked pycs: the source path recorded inside does not exist on the GPU box
            py_compile.compile(str(src), cfile=str(dst), dfile=f"<reference>/{name}.py", doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    (OUT / "PYTHON_VERSION").write_text(f"{sys.version_info[0]}.{sys.version_info[1]}\n")
    return True


def available() -> bool:
    tag = OUT / "PYTHON_VERSION"
    return (tag.exists() and tag.read_text().strip() == f"{sys.version_info[0]}.{sys.version_info[1]}" and
            all((OUT / f"{m}.pycode").exists() for m in MODULES))


class _Anything:
    """Stands in for every matplotlib object: any attribute, call, unpack or item works."""

    def __getattr__(self, attr):
        return self

    def __call__(self, *a, **k):
        return self

    def __iter__(self):
        return iter((self, self))


class _RefFinder:
    """Meta-path finder that serves the reference's module names from oracle/_ref/<name>.pycode
    (a .pyc by another name: 16-byte header + marshalled code object)."""

    def find_spec(self, name, path=None, target=None):
        import importlib.util

        if name in MODULES and (OUT / f"{name}.pycode").exists():
            return importlib.util.spec_from_loader(name, self, origin=str(OUT / f"{name}.pycode"))
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import marshal

        blob = (OUT / f"{module.__name__}.pycode").read_bytes()
        module.__file__ = str(OUT / f"{module.__name__}.pycode")
        exec(marshal.loads(blob[16:]), module.__dict__)


def load() -> types.SimpleNamespace:
    """Import the byte-compiled reference.  Its modules import each other by their plain names, so they are
    imported under those names through a temporary finder and then taken out of sys.modules again (the
    product package has same-named modules of its own); callers hold on to the returned namespace."""
    if not available():
        raise ImportError("oracle/_ref is not built (run oracle/build_ref.py where /root/reference exists)")
    import importlib

    stubbed = "matplotlib" not in sys.modules
    if stubbed:   # scan_mismtach / plotting import matplotlib for their figures; it is not installed here

        def _plt_attr(attr):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return _Anything()

        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        plt.__getattr__ = _plt_attr  # type: ignore[attr-defined]
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    saved = {m: sys.modules.pop(m) for m in MODULES if m in sys.modules}
    finder = _RefFinder()
    sys.meta_path.insert(0, finder)
    try:
        mods = {m: importlib.import_module(m) for m in MODULES if m != "io_fwm"}
        for m in mods.values():
            assert str(OUT) in str(getattr(m, "__file__", "")), f"{m.__name__} did not come from oracle/_ref"
    finally:
        sys.meta_path.remove(finder)
        for m in MODULES:
            sys.modules.pop(m, None)
        sys.modules.update(saved)
        if stubbed:   # the reference modules keep their own reference to the stub; nobody else should see it
            sys.modules.pop("matplotlib", None)
            sys.modules.pop("matplotlib.pyplot", None)
    return types.SimpleNamespace(**mods)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print(f"oracle/_ref: {'built' if ok else 'unavailable (no /root/reference and no previous build)'}")
