"""The C ABI (include/rbphd.h) against the built library: every declared symbol is exported, the
ctypes structs match the header's layout, and the library refuses to compute without a device.
No compute call is made here (runs without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rbphd.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rbphd_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from monorfs_b200 import build, capi
    build.build()
    return capi.load()


def test_header_declares_the_binding_list():
    from monorfs_b200 import capi
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    from monorfs_b200 import capi
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    for s in declared_symbols():
        assert hasattr(lib, s)


def test_no_torch_or_cuda_types_in_signatures():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for banned in ("cudaStream_t", "at::", "torch", "cuda_runtime", "CUstream"):
        assert banned not in text, banned


def test_struct_layouts_match_header():
    from monorfs_b200 import capi
    # rbphd_config: 4 int32 + (9+36+1+1+9+1+1+1+1+1+1+3+7) doubles
    assert C.sizeof(capi.RbphdConfig) == 16 + 8 * 72
    assert capi.RbphdConfig.R.offset == 16
    assert capi.RbphdConfig.measurer.offset == 16 + 8 * 65
    assert C.sizeof(capi.RbphdLimits) == 32
    from oracle import orc
    assert C.sizeof(orc.OrcConfig) == C.sizeof(capi.RbphdConfig)
    for (name, _), (oname, _) in zip(capi.RbphdConfig._fields_, orc.OrcConfig._fields_):
        assert name == oname
        assert getattr(capi.RbphdConfig, name).offset == getattr(orc.OrcConfig, oname).offset


def test_fails_loudly_without_a_device(lib):
    """No CPU fallback: without a CUDA device rbphd_new returns NULL and says why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from monorfs_b200 import capi, synth
    sc = synth.make_scene(2, 5, 3, seed=1)
    with pytest.raises(capi.RbphdError) as e:
        capi.Handle(sc.params, max_particles=2)
    assert e.value.code == capi.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "monorfs_b200")
    banned = re.compile(r"import\s+oracle|from\s+oracle|oracle/|oracle\\|liborc|rbphd_oracle|orc_[a-z]+\(")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not banned.search(text), os.path.join(dirpath, f)


def test_integration_doc_binds_every_export():
    """INTEGRATION.md shows the C# side of the boundary: every entry point of include/rbphd.h has its
    DllImport stub there, and the stubs name no symbol the header does not declare."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    bound = set(re.findall(r"\b(rbphd_[a-z0-9_]+)\s*\(", text))
    declared = set(declared_symbols())
    assert not (declared - bound), sorted(declared - bound)
    assert not (bound - declared), sorted(bound - declared)
