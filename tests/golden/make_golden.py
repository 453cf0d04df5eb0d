"""Generate the committed golden fixtures from the UNMODIFIED reference compiled in this container.

Run here (needs /root/reference and `make -C oracle ref`):  python tests/golden/make_golden.py
Outputs (committed):
  ve_kat.npz          inputs = sph/test/example_data.txt (99 x 31), reference J-loop outputs (oracle/_ref/ref_kat),
                      and the literal expectations + tolerances of sph/test/ve.cpp:112-233
  sedov12_step0.npz   \
  sedov12_step2.npz    | full per-stage dumps of oracle/_ref/ref_harness (reference CPU code, -ffp-contract=off):
  noh14_step0.npz      | inputs, tree view, nc, sorted CSR neighbour lists, every loop output, post-integrate state
  turb12_step0.npz     |
  turb12h_step0.npz   /  (hscale=1.6: exercises the h-iteration both ways)
  turb12av_step0.npz     the same with avClean = true (HydroVeProp<true>: dV11..dV33, computeMomentumEnergy<true>)
  sedov16_energies.npz  etot/ecin/eint series over 30 steps
  sedov50_energies.npz  the same over 100 steps of BASELINE config 0 (sedov -n 50)
  noh14_energies.npz, turb12_energies.npz  40 steps continuing noh14_step0 / turb12_step0 (open box that follows the
                        particles; periodic box with a velocity field)
  turb12s_step0.npz, turb12s_step2.npz  the reference's turbulence-ve propagator (ref_harness stir=1: particles at rest,
                        sph::driveTurbulence after the momentum loop): inputs, loop outputs, TurbulenceData state,
                        accelerations after stirring; no neighbour lists / tree (the test builds its own)
  turb12s_energies.npz  40 steps of the same
  turb_form2.npz        TurbulenceData state for stSpectForm = 2 (modes drawn from the random engine)
"""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from refdata import GOLDEN, REPO, csr_sorted_neighbors, run_ref_harness  # noqa: E402

REF = Path("/root/reference")


def make_kat():
    data = np.loadtxt(REF / "sph/test/example_data.txt")
    assert data.shape == (99, 31)
    out = subprocess.run([str(REPO / "oracle/_ref/ref_kat"), str(REF / "sph/test/example_data.txt")], check=True,
                         stdout=subprocess.PIPE, text=True).stdout
    ref = {}
    for line in out.splitlines():
        k, v = line.split()
        ref["ref_" + k] = np.float64(v)
    # literal expectations of sph/test/ve.cpp (value, abs tolerance)
    expect = {
        "av_alpha": (0.93941905320351171, 2e-9),
        "dc_divv": (3.3760353440920682e-2, 2e-9), "dc_curlv": (3.7836647734377962e-2, 2e-9),
        "dc_dV11": (0.0013578323369918166, 2e-9), "dc_dV12": (0.02465266861727711, 2e-9),
        "dc_dV13": (-0.0046604174274769167, 2e-9), "dc_dV22": (0.022556438947324862, 2e-9),
        "dc_dV23": (0.0097704904179710741, 2e-9), "dc_dV33": (0.0098460821566040066, 2e-9),
        "iad_0": (1.9296619855715329e-18, 1e-10), "iad_1": (-1.7838691836843698e-20, 1e-10),
        "iad_2": (-1.2892885646884301e-20, 1e-10), "iad_3": (1.9482845913025683e-18, 1e-10),
        "iad_4": (1.635410357476855e-20, 1e-10), "iad_5": (1.9246939006338132e-18, 1e-10),
        "mom1_ax": (-505548.68073726865, 0.023), "mom1_ay": (303384.91384746187, 0.053),
        "mom1_az": (-1767463.9739728321, 0.043), "mom1_du": (8.5525242525359648e12, 7.1e5),
        "mom1_maxvsignal": (26490876.319252387, 1e-6),
        "mom0_ax": (-521261.07791667967, 0.022), "mom0_ay": (-74471.016515749841, 0.064),
        "mom0_az": (-1730426.827721074, 0.042), "mom0_du": (7.1838438980436924e12, 3.1e5),
        "mom0_maxvsignal": (26490876.319252387, 1e-6),
        "gradh_density": (3.4662283566584293e1, 8e-7), "gradh_gradh": (0.98699067585409861, 5e-7),
        "gradh_kx": (1.0042661134076782, 3e-7),
        "xmass_rho0": (34.515038498081417, 7.33e-7),
    }
    exp = {"exp_" + k: np.array(v) for k, v in expect.items()}
    np.savez_compressed(GOLDEN / "ve_kat.npz", example_data=data, **ref, **exp)


def pack_step(d: dict) -> dict:
    d = dict(d)
    ngmax = int(d["ngmax"][0])
    off, idx = csr_sorted_neighbors(d.pop("neighbors"), d["nc"], ngmax)
    d["nb_offsets"] = off
    d["nb_sorted"] = idx
    d.pop("_step", None)
    return d


def make_steps():
    with tempfile.TemporaryDirectory() as tmp:
        for case, n, steps, keep, hs, av in [("sedov", 12, 3, (0, 2), 1.0, False), ("noh", 14, 1, (0,), 1.0, False),
                                             ("turb", 12, 1, (0,), 1.0, False), ("turb", 12, 1, (0,), 1.6, False),
                                             ("turb", 12, 1, (0,), 1.0, True)]:
            tag = f"{case}{n}" + ("" if hs == 1.0 else "h") + ("av" if av else "")
            dumps = run_ref_harness(case, n, steps, Path(tmp) / tag, hscale=hs, av_clean=av)
            for k in keep:
                d = pack_step(dumps[k])
                if not (case == "sedov" and k == 0):
                    d.pop("wh"), d.pop("whd")  # tables are identical in every dump; keep one copy
                np.savez_compressed(GOLDEN / f"{tag}_step{k}.npz", **d)
        out = Path(tmp) / "sedov16"
        run_ref_harness("sedov", 16, 30, out, dump_every=1000, dump_neighbors=False)
        e = np.loadtxt(out / "energies.txt")
        cols = np.array("step ttot minDt etot ecin eint linmom angmom totalNeighbors".split())
        np.savez_compressed(GOLDEN / "sedov16_energies.npz", series=e, columns=cols)
        for case, n, steps, tag in [("sedov", 50, 100, "sedov50"), ("noh", 14, 40, "noh14"), ("turb", 12, 40, "turb12")]:
            out = Path(tmp) / (tag + "_e")
            run_ref_harness(case, n, steps, out, dump_every=100000, dump_neighbors=False)
            np.savez_compressed(GOLDEN / f"{tag}_energies.npz", series=np.loadtxt(out / "energies.txt"), columns=cols)


def make_stirring():
    cols = np.array("step ttot minDt etot ecin eint linmom angmom totalNeighbors".split())
    with tempfile.TemporaryDirectory() as tmp:
        dumps = run_ref_harness("turb", 12, 3, Path(tmp) / "s", dump_neighbors=False, stir=1)
        for k in (0, 2):
            d = {key: v for key, v in dumps[k].items() if not key.startswith("tree_") and key not in ("wh", "whd", "_step")}
            np.savez_compressed(GOLDEN / f"turb12s_step{k}.npz", **d)
        out = Path(tmp) / "e"
        run_ref_harness("turb", 12, 40, out, dump_every=100000, dump_neighbors=False, stir=1)
        np.savez_compressed(GOLDEN / "turb12s_energies.npz", series=np.loadtxt(out / "energies.txt"), columns=cols)
        d = run_ref_harness("turb", 6, 1, Path(tmp) / "f2", dump_neighbors=False, stir=2)[0]
        np.savez_compressed(GOLDEN / "turb_form2.npz", **{k: v for k, v in d.items() if k.startswith("turb_")})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stirring":
        make_stirring()
    else:
        make_kat()
        make_steps()
        make_stirring()
    for f in sorted(GOLDEN.glob("*.npz")):
        print(f.name, f.stat().st_size)
