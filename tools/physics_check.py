"""Statistical parity of the HMC chain (plaquette, acceptance, <exp(-dH)>) at BASELINE config 1:
64x64, beta=2, m0=0, MD=10, tau=1.

    python tools/physics_check.py gpu  [ntherm nmeas]      chain on the GPU (device RNG, host Metropolis)
    python tools/physics_check.py gpu-evenodd [ntherm nmeas]   the same with the opt-in even-odd HMC (another Markov chain, same distribution)
    python tools/physics_check.py cpu  [ntherm nmeas nchains]   the same chain with the C oracle on CPU cores
Both print one JSON line; profiles/r01_physics_64x64.json holds the pair."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.port import Port  # noqa: E402

NX = NT = 64
BETA, M0, MD, TAU = 2.0, 0.0, 10, 1.0


def binned_error(x, nbins=10):
    x = np.asarray(x, float)
    n = len(x) // nbins * nbins
    b = x[:n].reshape(nbins, -1).mean(axis=1)
    return float(b.std(ddof=1) / np.sqrt(nbins))


def summarize(plaq, acc, dh):
    return {"plaquette": float(np.mean(plaq)), "plaquette_err": binned_error(plaq), "acceptance": float(np.mean(acc)),
            "acceptance_err": binned_error(np.asarray(acc, float)), "exp_minus_dH": float(np.mean(np.exp(-np.asarray(dh)))),
            "exp_minus_dH_err": binned_error(np.exp(-np.asarray(dh))), "mean_dH": float(np.mean(dh)), "n": len(plaq)}


def cpu_chain(args):
    seed, ntherm, nmeas = args
    P = Port(NX, NT)
    rng = np.random.default_rng(seed)
    U = P.hot_start(1000 + seed)
    V = NX * NT
    plaq, acc, dh = [], [], []
    for i in range(ntherm + nmeas):
        pi = rng.normal(size=(2, V))
        chi = (rng.normal(size=(2, V)) + 1j * rng.normal(size=(2, V))) / np.sqrt(2.0)
        t = P.trajectory(U, pi, chi, MD, TAU, BETA, M0)
        a = rng.random() <= np.exp(-t["dH"]) if t["dH"] > -700 else True
        if a:
            U = t["U"]
        if i >= ntherm:
            plaq.append(P.plaquette(U, BETA)[1] / V)
            acc.append(bool(a))
            dh.append(t["dH"])
    return plaq, acc, dh


def main():
    mode = sys.argv[1]
    ntherm = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    nmeas = int(sys.argv[3]) if len(sys.argv) > 3 else 600
    t0 = time.time()
    if mode in ("gpu", "gpu-evenodd"):
        import schwingermodel_b200 as sb
        lat = sb.Lattice(NX, NT)
        h = sb.HMC(lat, Port(NX, NT).hot_start(12345), MD, TAU, ntherm, nmeas, 0, BETA, M0, seed=2024)
        # thermalisation always with the reference-exact solver: from the hot start the even-odd action's leapfrog error at
        # eps = 0.1 is dH ~ 12 on 64^2 (reference action: ~5) and the chain would never accept; from a thermalised field
        # its dH is the smaller of the two (profiles/r02_evenodd_dH_scaling.txt)
        for _ in range(ntherm):
            h.HMC_Update()
        h.therm = True
        if mode == "gpu-evenodd":
            lat.set_solver("evenodd")
        plaq = []
        for _ in range(nmeas):
            h.HMC_Update()
            plaq.append(h.sum_re_plaq / (NX * NT))
        hist = h.history[ntherm:]
        res = summarize(plaq, [x[1] for x in hist], [x[0] for x in hist])
        res.update(impl="b200" if mode == "gpu" else "b200, opt-in even-odd HMC (thermalised with the reference solver)",
                   seconds=time.time() - t0, ntherm=ntherm)
    else:
        from multiprocessing import Pool
        nchains = int(sys.argv[4]) if len(sys.argv) > 4 else 6
        with Pool(nchains) as pool:
            outs = pool.map(cpu_chain, [(s, ntherm, nmeas) for s in range(nchains)])
        plaq = sum((o[0] for o in outs), [])
        acc = sum((o[1] for o in outs), [])
        dh = sum((o[2] for o in outs), [])
        res = summarize(plaq, acc, dh)
        res.update(impl="oracle (C restatement of the reference, bit-exact to it)", seconds=time.time() - t0,
                   ntherm=ntherm, chains=nchains)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
