"""The drop-in driver's front end on the CPU: the reference's three input files are parsed
(kernel.cu:542-604, :693-700), model and operator are prepared, and -- without a GPU -- the run stops
with the no-device error instead of falling back to anything."""
import os

import pytest

import rtm_gpu_b200 as R
from golden_cases import GOLDEN_CASES
from refcase import data_tiny, velocity_tiny, write_inputs

RTM_ERR_IO, RTM_ERR_NO_DEVICE = -4, -6


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
