"""Build the UNMODIFIED reference rasterizer into oracle/_ref/ (test/bench infrastructure only).

TEST INFRASTRUCTURE -- nothing under generativedensification_b200/ or
diff_gaussian_rasterization/ may import anything from oracle/.

The reference's rasterizer (third_party/diff-gaussian-rasterization in the
reference tree) is a small CUDA/C++ extension: ext.cpp, rasterize_points.cu and
cuda_rasterizer/{rasterizer_impl,forward,backward}.cu plus vendored GLM headers.
This recipe compiles those sources *where they lie* under /root/reference with
nvcc (through torch.utils.cpp_extension.load, i.e. ninja + nvcc; the reference's
own setup.py is not run) and writes only into oracle/_ref/:

    oracle/_ref/diff_gaussian_rasterization/_C.so      compiled extension (sm_100)
    oracle/_ref/diff_gaussian_rasterization/__init__.py installed copy of the
        reference's Python API file (what `pip install --target` would place
        there) so the reference can be driven through its own public API on the
        GPU box, where /root/reference does not exist.

oracle/_ref/ is git-ignored (never enters history) but not gpurun-ignored, so the
built files travel to the GPU box.  The reference needs a GPU to *run*; building
needs none.

The compile flags mirror the reference's setup.py:21-29: default nvcc flags (no
--use_fast_math), GLM include path, arch from TORCH_CUDA_ARCH_LIST=10.0.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("GDR_REFERENCE_ROOT", "/root/reference")
RAST = os.path.join(REF_ROOT, "third_party", "diff-gaussian-rasterization")
OUT = os.path.join(HERE, "_ref", "diff_gaussian_rasterization")


def have_reference_sources() -> bool:
    return os.path.isfile(os.path.join(RAST, "rasterize_points.cu"))


def is_built() -> bool:
    return os.path.isfile(os.path.join(OUT, "_C.so")) and os.path.isfile(os.path.join(OUT, "__init__.py"))


def build(verbose: bool = False) -> str:
    """Compile the reference extension; returns the output directory."""
    if not have_reference_sources():
        if is_built():
            return OUT
        raise RuntimeError(f"reference sources not found under {RAST} and no prebuilt oracle/_ref")
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load

    sources = [
        os.path.join(RAST, "cuda_rasterizer", "rasterizer_impl.cu"),
        os.path.join(RAST, "cuda_rasterizer", "forward.cu"),
        os.path.join(RAST, "cuda_rasterizer", "backward.cu"),
        os.path.join(RAST, "rasterize_points.cu"),
        os.path.join(RAST, "ext.cpp"),
    ]
    load(
        name="_C",
        sources=sources,
        extra_include_paths=[os.path.join(RAST, "third_party", "glm")],
        extra_cuda_cflags=["-lineinfo"],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=False,  # just build; importing happens in oracle/ref_api.py
    )
    shutil.copyfile(
        os.path.join(RAST, "diff_gaussian_rasterization", "__init__.py"),
        os.path.join(OUT, "__init__.py"),
    )
    # ninja leaves object files behind; keep the directory small for gpurun snapshots
    for f in os.listdir(OUT):
        if f.endswith(".o") or f.startswith(".ninja") or f == "build.ninja":
            try:
                os.remove(os.path.join(OUT, f))
            except OSError:
                pass
    return OUT


if __name__ == "__main__":
    out = build(verbose="-v" in sys.argv)
    print("reference built into", out, os.listdir(out))
