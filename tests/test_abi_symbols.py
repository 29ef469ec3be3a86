"""The C-ABI library loads without a GPU and exports every symbol include/fmcmc_b200.h declares."""
import ctypes as C
import os
import re

import fmcmc_b200
from fmcmc_b200 import _abi as A
from fmcmc_b200._lib import EXPORTED_SYMBOLS, SO_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "fmcmc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fmcmc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    fmcmc_b200.build()
    L = C.CDLL(SO_PATH)
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(EXPORTED_SYMBOLS) == names


def test_pure_host_entry_points():
    L = fmcmc_b200.lib()
    assert L.fmcmc_version() == A.ABI_VERSION
    assert L.fmcmc_device_count() >= 0
    assert L.fmcmc_rows_kept(500, 100, 7) == 57 == A.rows_kept(500, 100, 7)
    assert L.fmcmc_rows_kept(10, 0, 1) == 10
    assert L.fmcmc_kernel_state_len(A.KERNEL_ADAPT, 5, 4) == 20 == A.state_len(A.KERNEL_ADAPT, 5, 4)
    assert L.fmcmc_kernel_state_len(A.KERNEL_RAM, 5, 4) == 16
    assert L.fmcmc_kernel_state_len(A.KERNEL_NMIRROR, 5, 4) == 15
    assert L.fmcmc_kernel_state_len(A.KERNEL_NORMAL, 5, 4) == 0
    d = A.marshal_model(A.FAMILY_GAUSSIAN_LM, 10, p_x=2, flags=A.MODEL_INTERCEPT)
    assert L.fmcmc_model_nparams(d.byref()) == 4
    d = A.marshal_model(A.FAMILY_HIER_NORMAL, 10, n_groups=20, flags=A.MODEL_SCALES)
    assert L.fmcmc_model_nparams(d.byref()) == 23


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors == what a C compiler makes of the header."""
    import subprocess
    import tempfile
    src = ('#include "%s/include/fmcmc_b200.h"\n#include <stdio.h>\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n",'
           'sizeof(fmcmc_model_desc),sizeof(fmcmc_kernel_spec),sizeof(fmcmc_kernel_state),sizeof(fmcmc_stream_spec),'
           'sizeof(fmcmc_run_spec),sizeof(fmcmc_run_report),sizeof(fmcmc_shard_handles));return 0;}') % ROOT
    with tempfile.TemporaryDirectory() as td:
        c, exe = os.path.join(td, "s.c"), os.path.join(td, "s")
        open(c, "w").write(src)
        subprocess.run(["gcc", "-o", exe, c], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(A.ModelDesc), C.sizeof(A.KernelSpec), C.sizeof(A.KernelState), C.sizeof(A.StreamSpec),
                     C.sizeof(A.RunSpec), C.sizeof(A.RunReport), C.sizeof(A.ShardHandles)]


def test_no_gpu_is_a_loud_error():
    """Without a CUDA device the product fails with FMCMC_ECUDA: there is no CPU fallback."""
    import numpy as np
    import pytest
    if fmcmc_b200.lib().fmcmc_device_count() > 0:
        pytest.skip("a GPU is present")
    fam = fmcmc_b200.ll_gaussian_lm(np.zeros(4), np.zeros(4))
    with pytest.raises(fmcmc_b200.FmcmcError) as ei:
        fmcmc_b200.MCMC([0, 0, 1.0], fam, 10)
    assert ei.value.code == A.ECUDA
