"""The drop-in driver's front end on the CPU: the reference's three input files are parsed
(kernel.cu:542-604, :693-700), model and operator are prepared, and -- without a GPU -- the run stops
with the no-device error instead of falling back to anything."""
import os

import pytest

import rtm_gpu_b200 as R
from golden_cases import GOLDEN_CASES
from refcase import data_tiny, velocity_tiny, write_inputs

RTM_ERR_ARG, RTM_ERR_IO, RTM_ERR_NO_DEVICE = -1, -4, -6


def drive(wd, run_file="2D_Real_RVSP_RTM.txt", verbose=0):
    L = R.lib()
    cwd = os.getcwd()
    os.chdir(wd)
    try:
        rc = L.rtm_run_driver(run_file.encode(), 0, 0, verbose)
    finally:
        os.chdir(cwd)
    return rc, L.rtm_last_error().decode(errors="replace")


def test_missing_and_truncated_run_files(tmp_path):
    rc, err = drive(tmp_path)
    assert rc == RTM_ERR_IO and "cannot open run file" in err
    (tmp_path / "2D_Real_RVSP_RTM.txt").write_text("label\r\n10\r\nlabel\r\n2\r\n")
    rc, err = drive(tmp_path)
    assert rc == RTM_ERR_IO and "fewer than 28 values" in err


@pytest.mark.parametrize("crlf", [True, False])
def test_missing_parameter_and_model_files(tmp_path, crlf):
    case = GOLDEN_CASES["tiny_te_compen"]
    write_inputs(case, tmp_path, velocity_tiny(case), {d: data_tiny(case, d) for d in case.depths}, crlf=crlf)
    os.rename(tmp_path / "in" / "vel.dat", tmp_path / "in" / "vel.moved")
    rc, err = drive(tmp_path)
    assert rc == RTM_ERR_IO and "vel.dat" in err
    os.rename(tmp_path / "in" / "vel.moved", tmp_path / "in" / "vel.dat")
    os.remove(tmp_path / "in" / "Parameter.txt")
    rc, err = drive(tmp_path)
    assert rc == RTM_ERR_IO and "Parameter.txt" in err


def test_valid_inputs_stop_at_the_device_check_without_a_gpu(tmp_path, capfd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    case = GOLDEN_CASES["tiny_te_compen"]
    write_inputs(case, tmp_path, velocity_tiny(case), {d: data_tiny(case, d) for d in case.depths})
    rc, err = drive(tmp_path, verbose=1)
    out = capfd.readouterr().out
    assert rc == RTM_ERR_NO_DEVICE and "no CUDA device" in err
    # the echo of the derived sizes got as far as the velocity bins (kernel.cu:704-738)
    assert "nvel=" in out and "vmin=" in out


@pytest.mark.parametrize("field,value,needle", [("nrec", -3, "nrec"), ("n", 0, "n ="), ("NT1", -5, "NT1"), ("mod_NX", 0, "mod_NX"),
                                                ("ds", 0, "ds"), ("h", -20.0, "h =")])
def test_non_positive_sizes_are_argument_errors_not_crashes(tmp_path, field, value, needle):
    """nrec = -3 used to leave rtm_run_driver as an uncaught std::length_error."""
    import dataclasses
    case = GOLDEN_CASES["tiny_te_compen"]
    write_inputs(case, tmp_path, velocity_tiny(case), {d: data_tiny(case, d) for d in case.depths})
    bad = dataclasses.replace(case, **{field: value})
    par = [bad.h, bad.tao1, bad.mod_NZ, bad.mod_NX, bad.NT1, bad.s_l, bad.s_z, bad.n, bad.ds, bad.r_x, bad.nrec, bad.dr]
    (tmp_path / "in" / "Parameter.txt").write_text(" \n".join(("%.9g" % p) if isinstance(p, float) else str(p) for p in par) + "\n")
    rc, err = drive(tmp_path)
    assert rc == RTM_ERR_ARG and needle in err and "positive" in err, (rc, err)


def test_output_window_outside_the_model_is_rejected(tmp_path):
    import dataclasses
    case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], NX_ED=500)
    write_inputs(case, tmp_path, velocity_tiny(case), {d: data_tiny(case, d) for d in case.depths})
    rc, err = drive(tmp_path)
    assert rc == RTM_ERR_ARG and "output window" in err, (rc, err)


def test_memory_estimate_matches_the_layout():
    import ctypes as C
    p = R.Params(751, 2301, 10, 4, 7501, 1, 1, 4.0, 4.0, 4e-4, 20.0, 1e-4, 10, 12, 2301, 1, 1, 0)
    fixed, per = C.c_size_t(), C.c_size_t()
    assert R.lib().rtm_memory_estimate(C.byref(p), 7501, C.byref(fixed), C.byref(per)) == 0
    pitch = (22 + 2321 + 4 + 31) // 32 * 32          # padL = 22 for N2 = 10
    fields = 12 * 771 * pitch * 4
    strips = 2 * 7501 * 4 * (2301 + 751) * 4
    traces = 2 * 7501 * 2301 * 4
    assert per.value == fields + strips + traces + 2 * 2301 * 751 * 4 + 64
    assert 0.85e9 < per.value < 1.0e9 and fixed.value < 400e6   # C2 is ~0.9 GB per shot
