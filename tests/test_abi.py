"""CPU tests of the boundary: the shared library loads without a GPU, exports every
symbol include/kpp_gpu.h declares, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from mckpp_f90_b200 import capi, synth, build as kbuild, driver
from mckpp_f90_b200.fields import field_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    kbuild.build()
    return capi.load()


def test_library_exports_every_declared_symbol(lib):
    declared = capi.header_exports()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.kpp_gpu_abi_version() == 1


def test_library_is_built_for_sm_100a():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True)
    assert "sm_100a" in out.stdout


def test_field_table_matches_header_and_data_model(lib):
    ids = capi.FIELD_IDS
    assert ids["KPP_F_U"] == 0 and "KPP_F__COUNT" in ids
    shapes = field_shapes(synth.make_case(synth.scaled(synth.CONFIGS["cfg1"], 2, 2))[0].dims)
    for cname, fid in ids.items():
        if cname == "KPP_F__COUNT":
            continue
        pyname = lib.kpp_gpu_field_name(fid).decode()
        assert pyname.lower().replace("_", "") == cname[6:].lower().replace("_", "").replace("diag", "diag"), cname
        if not pyname.startswith("diag_"):
            assert pyname in shapes, pyname
    for n in driver.INPUT_FIELDS + driver.ALL_OUTPUTS:
        assert n in capi.FIELD_BY_NAME, n


def test_no_cpu_fallback_without_device(lib):
    if lib.kpp_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cf, f, r = synth.make_case(synth.scaled(synth.CONFIGS["cfg1"], 2, 2))
    with pytest.raises(capi.KppError) as e:
        capi.KppGpu(cf)
    assert e.value.code == capi.KPP_E_NODEVICE
    with pytest.raises(capi.KppError):
        capi.test_eos(np.array([35.0]), np.array([10.0]), np.array([100.0]))
    assert b"no CPU fallback" in lib.kpp_gpu_strerror(capi.KPP_E_NODEVICE)


def test_argument_validation_happens_before_device_use(lib):
    cf, f, r = synth.make_case(synth.scaled(synth.CONFIGS["cfg1"], 2, 2))
    cf.consts.LKPP = False
    with pytest.raises(capi.KppError) as e:
        capi.KppGpu(cf)
    assert e.value.code == capi.KPP_E_INVALID


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "mckpp_f90_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h", ".cpp", ".F90")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle_lib" not in txt and "libmckpp_oracle" not in txt and "mckpp_oracle.h" not in txt, fn


def test_fortran_shim_field_ids_match_header():
    """The ISO_C_BINDING shim (delivered as source; no Fortran compiler here) hard-codes the
    enum values of include/kpp_gpu.h and the BIND(C) struct layouts: keep them in step."""
    src = open(os.path.join(ROOT, "mckpp_f90_b200", "fortran", "mckpp_physics_driver_gpu.F90")).read()
    pairs = dict((k, int(v)) for k, v in re.findall(r"(KPP_F_\w+)\s*=\s*(\d+)", src))
    assert len(pairs) >= 59
    for name, val in pairs.items():
        assert capi.FIELD_IDS[name] == val, name
    outs = dict((k, int(v)) for k, v in re.findall(r"(KPP_OUT_\w+)\s*=\s*(\d+)", src))
    want = {k: v for k, v in capi.out_ids().items() if k != "KPP_OUT__COUNT"}
    assert outs == want
    # struct members in the same order as the C structs
    consts_members = re.search(r"TYPE, BIND\(C\) :: kpp_consts(.*?)END TYPE kpp_consts", src, flags=re.S).group(1)
    order = re.findall(r"\b(dto|grav|vonk|sice|hmixtolfrac|iso_thresh|itermax|iso_bot|dt_uvdamp|LKPP|LRI|LDD|L_SSref|"
                       r"L_RELAX_SST|L_RELAX_CALCONLY|L_FCORR|L_FCORR_WITHZ|L_SFCORR|L_SFCORR_WITHZ|L_RELAX_SAL|"
                       r"L_RELAX_OCNT|L_NO_FREEZE|L_NO_ISOTHERM|L_DAMP_CURR|L_VARY_BOTTOM_TEMP|have_ocnT_file|"
                       r"have_sal_file|numerics|reserved)\b", consts_members)
    assert order == capi._CONST_D + capi._CONST_I
    hdr = open(capi.HEADER_PATH).read()
    cstruct = re.search(r"typedef struct kpp_consts \{(.*?)\} kpp_consts;", hdr, flags=re.S).group(1)
    cstruct = re.sub(r"/\*.*?\*/", "", cstruct, flags=re.S)
    corder = re.findall(r"\b(\w+)\s*[,;]", cstruct)
    assert corder == capi._CONST_D + capi._CONST_I


def _shim_source():
    return open(os.path.join(ROOT, "mckpp_f90_b200", "fortran", "mckpp_physics_driver_gpu.F90")).read()


def test_fortran_shim_binds_only_declared_entry_points():
    """Every BIND(C, name=...) of the shim is an entry point include/kpp_gpu.h declares (and the library exports)."""
    hdr = open(capi.HEADER_PATH).read()
    declared = set(re.findall(r"\b(kpp_gpu_\w+)\s*\(", hdr))
    bound = set(re.findall(r'BIND\(C,\s*name="(\w+)"\)', _shim_source())) - {"strlen"}      # libc, for the error text
    assert len(bound) >= 15
    assert bound <= declared, bound - declared


def test_fortran_shim_checks_every_return_code():
    """no return code of the library may be dropped.  Every reference to an INTEGER(c_int) entry point in
    the executable part is the argument of `check(...)`, or assigns `rc` that a later `check(rc, ...)` tests;
    kpp_gpu_destroy at finalisation is the one exception."""
    src = _shim_source()
    body = src[src.index("CONTAINS"):]
    body = "\n".join(ln.split("!")[0] for ln in body.splitlines())        # strip comments
    body = re.sub(r"&\s*\n\s*", " ", body)                                 # join continuation lines
    int_funcs = set(re.findall(r"INTEGER\(c_int\) FUNCTION (kpp_gpu_\w+)", src))
    assert "kpp_gpu_step" in int_funcs and "kpp_gpu_sync" in int_funcs
    lines = body.splitlines()
    for i, ln in enumerate(lines):
        for fn in re.findall(r"\b(kpp_gpu_\w+)\s*\(", ln):
            if fn not in int_funcs or fn == "kpp_gpu_destroy":
                continue
            if re.search(r"CALL\s+check\(\s*" + fn + r"\b", ln):
                continue
            m = re.match(r"\s*rc\s*=\s*" + fn + r"\b", ln)
            assert m, f"return code of {fn} dropped: {ln.strip()}"
            rest = "\n".join(lines[i + 1:i + 40])
            assert re.search(r"CALL\s+check\(\s*rc\b", rest), f"rc of {fn} never checked"


def test_fortran_shim_assumed_rank_only_inside_select_rank():
    """ADVICE r1 (F2018 C839): an assumed-rank dummy may only be a SELECT RANK selector (or go to an inquiry
    intrinsic); it must never be passed on as an actual argument."""
    src = _shim_source()
    for m in re.finditer(r"SUBROUTINE (\w+)\((.*?)\)\n(.*?)END SUBROUTINE \1", src, flags=re.S):
        text = m.group(3)
        for name in re.findall(r"::\s*(\w+)\(\.\.\)", text):
            code = "\n".join(ln.split("!")[0] for ln in text.splitlines())
            code = re.sub(r"::\s*" + name + r"\(\.\.\)", "", code)
            outside = re.sub(r"SELECT RANK \(" + name + r"\).*?END SELECT", "", code, flags=re.S)
            assert not re.search(r"\b" + name + r"\b", outside), (m.group(1), name)
            assert re.search(r"SELECT RANK \(" + name + r"\)", code), (m.group(1), name)
