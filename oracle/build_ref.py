"""Byte-compile the reference's own hot-path modules, from the sources where they lie, into
oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot like the built .so files).

    python oracle/build_ref.py            # build container only: needs /root/reference

The reference is pure Python (no build system, not pip-installable: no setup.py / pyproject.toml), so
"building" it is `py_compile`.  Nothing of its source text enters the repository; the .pyc files are
build outputs.  `load()` imports them with the stub recipe of SURVEY.md section 8(c) (imageio,
configargparse and skimage are absent from the image and unused by the path; `batchify_rays`, the
undefined name render() calls at main.py:79, is aliased to the function that exists, batch_rays).

TEST INFRASTRUCTURE ONLY: used by bench.py's `--impl reference` arm / cpu_baseline and by tests.
"""
import marshal
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference"
MODULES = ("utils", "model", "data_helpers", "main")      # import order of the reference (main.py:1-13)


def build():
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    for name in MODULES:      # ".bytecode", not ".pyc": snapshot tools tend to drop *.pyc as cache files
        py_compile.compile(os.path.join(REF, name + ".py"), cfile=os.path.join(OUT, name + EXT), doraise=True,
                           dfile=f"/root/reference/{name}.py")
    return True


EXT = ".bytecode"


def available():
    return all(os.path.exists(os.path.join(OUT, m + EXT)) for m in MODULES)


def load():
    """-> the reference's `main` module (with `model`, `utils`, `data_helpers` importable), run from
    the byte-compiled files.  The modules are registered under their own top-level names because the
    reference imports them that way (`from model import *`)."""
    import torch
    for name in ["imageio", "configargparse", "skimage", "skimage.transform", "skimage.io", "skimage.color"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage.transform"].rescale = None
    sys.modules["skimage.io"].imread = None
    sys.modules["skimage.color"].rgba2rgb = None
    loaded = {}
    for name in MODULES:
        if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(OUT):
            loaded[name] = sys.modules[name]
            continue
        path = os.path.join(OUT, name + EXT)
        with open(path, "rb") as fh:
            code = marshal.loads(fh.read()[16:])          # 16-byte pyc header, then the code object
        mod = types.ModuleType(name)
        mod.__file__ = path
        sys.modules[name] = mod
        exec(code, mod.__dict__)
        loaded[name] = mod
    ref_main = loaded["main"]
    ref_main.batchify_rays = ref_main.batch_rays
    # main.py:15 picks "cuda" when a GPU is visible; this is the CPU baseline, so its models stay on the
    # host like its inputs (on the GPU box the module would otherwise mix devices)
    ref_main.device = torch.device("cpu")
    torch.autograd.set_detect_anomaly(False)      # main.py:16 switches it on at import; irrelevant to a no_grad render
    return ref_main


if __name__ == "__main__":
    print("oracle/_ref built" if build() else "no /root/reference here: nothing built")
