#!/usr/bin/env python
"""Generates tests/golden/ref_*.npz by RUNNING THE REFERENCE'S OWN SOURCES (oracle/_ref/libpinc_ref.so, compiled
in place from /root/reference/src by oracle/Makefile) on seeded inputs.  Only works where /root/reference exists;
the fixtures it writes are committed and are what pins the oracle on boxes without the reference.

    python tests/golden/make_ref_fixtures.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import small_cfg  # noqa: E402
from oracle import ref  # noqa: E402
from pinc_b200 import abi, initial  # noqa: E402

SCENARIOS = {
    # BASELINE config 1 scaled down: lattice + perturbation over 1,2,2 sub-domains
    "cold": dict(base="cold", steps=6, ic="lattice",
                 over=dict(grid__truesize="16,8,8", multigrid__mglevels=3, population__nparticles="8 pc",
                           population__nalloc="16 pc", population__perturbamplitude="2e-3,0,0,0,0,0",
                           grid__nemigrantsalloc="4 pc")),
    # BASELINE config 2 scaled down: Maxwellian plasma on one sub-domain (self-migration through 26 neighbours)
    "warm": dict(base="warm", steps=6, ic="maxwellian",
                 over=dict(grid__truesize="16,16,16", multigrid__mglevels=3, population__nparticles="8 pc",
                           population__nalloc="16 pc", population__thermalvelocitycells="0.08,0.004",
                           grid__nemigrantsalloc="2 pc")),
    # BASELINE config 4 scaled down: warm plasma over 1,2,2 sub-domains (real cross-rank migration)
    "warm4": dict(base="warm_big", steps=5, ic="maxwellian",
                  over=dict(grid__truesize="16,8,8", multigrid__mglevels=3, population__nparticles="8 pc",
                            population__nalloc="16 pc", population__thermalvelocitycells="0.08,0.004",
                            grid__nemigrantsalloc="4 pc")),
}


def initial_particles(cfg, ic):
    if ic == "lattice":
        return initial.perturb(cfg, initial.lattice(cfg))
    return initial.maxwellian(cfg, seed=7)


def moments(world, cfg):
    """Per rank and species: count, sum and sum of squares of every phase-space coordinate (order independent)."""
    out = []
    for r in range(cfg.nRanks):
        for s, (p, v) in enumerate(world.particles(r)):
            rec = np.concatenate([p, v], axis=1)
            out.append(np.concatenate([[len(p)], rec.sum(0), (rec * rec).sum(0)]))
    return np.array(out)


def run_world(world, cfg, steps, ic):
    world.set_particles(initial_particles(cfg, ic))
    world.migrate(); world.field_solve(); world.half_kick()
    rec = {"moments": [], "energy": [], "nEmig": [], "nImm": []}
    for it in range(steps):
        world.step()
        rec["moments"].append(moments(world, cfg))
        rec["energy"].append(world.energies())
    out = {k: np.array(v) for k, v in rec.items() if v}
    for name in ("rho", "phi", "E"):
        out[name] = np.stack([np.asarray(world.grid(r, name)).reshape(-1) for r in range(cfg.nRanks)])
    return out


def main():
    assert ref.available(), "oracle/_ref/libpinc_ref.so missing: make -C oracle ref (needs /root/reference)"
    for name, sc in SCENARIOS.items():
        text, cfg = small_cfg(sc["base"], **sc["over"])
        W = ref.RefWorld(text, cfg.nRanks)
        out = run_world(W, cfg, sc["steps"], sc["ic"])
        out["nEmig_last"] = np.array([[W.ranks[r].mpi.contents.nEmigrants[i] for i in range(27 * cfg.nSpecies)] for r in range(cfg.nRanks)])
        out["charge"] = np.array(cfg.charge); out["mass"] = np.array(cfg.mass)
        W.close()
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
