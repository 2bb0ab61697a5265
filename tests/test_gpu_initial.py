"""Initial conditions on the device against the host restatement of pPosLattice/pPosPerturb
(pinc_b200/initial.py, the same arithmetic as src/population.c:172-276) and, for the random starts, through
properties: the particle set does not depend on the decomposition, every particle lies in its rank's sub-domain,
seeds reproduce, Maxwellian moments."""
import numpy as np
import pytest

from helpers import da, ia, la, small_cfg, sorted_particles
from pinc_b200 import abi, initial

pytestmark = pytest.mark.gpu


def rank_structs(L, cfg, r):
    m = L.pincMpiAlloc(3, cfg.nSpecies, ia(cfg.nSubdomains), ia(cfg.nGhostLayers), ia(cfg.trueSize), r, cfg.nRanks)
    per_rank = [-(-a // cfg.nRanks) for a in cfg.nAlloc]
    p = L.pincPopAlloc(cfg.nSpecies, 3, la(per_rank), da(cfg.charge), da(cfg.mass))
    return m, p


def fetch(L, p):
    L.pincSyncPopToHost(p)
    pc = p.contents
    pos, vel = abi.pop_arrays(pc)
    return [(pos[pc.iStart[s]:pc.iStop[s]].copy(), vel[pc.iStart[s]:pc.iStop[s]].copy()) for s in range(pc.nSpecies)]


def test_lattice_and_perturbation_match_host(gpu_lib):
    L = gpu_lib
    text, cfg = small_cfg("cold", grid__truesize="16,8,8", population__nparticles="8 pc", population__nalloc="16 pc",
                          population__perturbamplitude="2e-3,0,0,0,1e-3,0")
    ref = initial.perturb(cfg, initial.lattice(cfg))
    for r in range(cfg.nRanks):
        m, p = rank_structs(L, cfg, r)
        L.pincPosLattice(p, m, la(cfg.nParticles), ia(cfg.trueSize))
        L.pincPosPerturb(p, m, da(cfg.perturbAmplitude), da(cfg.perturbMode), ia(cfg.trueSize))
        L.pincVelZero(p)
        got = fetch(L, p)
        for s in range(cfg.nSpecies):
            assert len(got[s][0]) == len(ref[r][s][0])
            a, b = sorted_particles(*got[s]), sorted_particles(*ref[r][s])
            assert np.abs(a - b).max() <= 1e-12          # device pow/fmod/cos differ from libm by ulps
            assert not got[s][1].any()
        L.pincPopFree(p); L.pincMpiFree(m)


def test_uniform_maxwellian_properties(gpu_lib):
    L = gpu_lib
    text4, cfg4 = small_cfg("warm_big", grid__nsubdomains="1,2,2", grid__truesize="16,8,8", population__nparticles="16 pc",
                            population__nalloc="48 pc", population__thermalvelocitycells="0.05,0.002")
    text1, cfg1 = small_cfg("warm_big", grid__nsubdomains="1,1,1", grid__truesize="16,16,16", population__nparticles="16 pc",
                            population__nalloc="48 pc", population__thermalvelocitycells="0.05,0.002")
    assert cfg1.nParticles == cfg4.nParticles

    def generate(cfg, seed):
        out = []
        for r in range(cfg.nRanks):
            m, p = rank_structs(L, cfg, r)
            L.pincPosUniform(p, m, la(cfg.nParticles), ia(cfg.trueSize), seed)
            L.pincVelMaxwell(p, m, da(cfg.drift), da(cfg.thermalVelocity), seed + 1)
            parts = fetch(L, p)
            off = np.array(initial.rank_offset(r, cfg), dtype=float)
            sub = np.array(initial.rank_subdomain(r, cfg.nSubdomains))
            for s in range(cfg.nSpecies):
                g = parts[s][0] + off
                lo = sub * np.array(cfg.trueSize)
                assert (g >= lo).all() and (g < lo + np.array(cfg.trueSize)).all()      # inside the rank's sub-domain
            out.append([(parts[s][0] + off, parts[s][1]) for s in range(cfg.nSpecies)])
            L.pincPopFree(p); L.pincMpiFree(m)
        return out

    one, four, again, other = generate(cfg1, 11), generate(cfg4, 11), generate(cfg4, 11), generate(cfg4, 12)
    for s in range(2):
        all4 = np.concatenate([four[r][s][0] for r in range(4)])
        assert len(all4) == cfg4.nParticles[s] == len(one[0][s][0])                    # nobody lost, nobody duplicated
        a = all4[np.lexsort(all4.T[::-1])]
        b = one[0][s][0][np.lexsort(one[0][s][0].T[::-1])]
        assert np.abs(a - b).max() <= 1e-12                                              # same set for any decomposition
        ag = np.concatenate([again[r][s][0] for r in range(4)])
        assert np.array_equal(np.sort(ag, axis=0), np.sort(all4, axis=0))                # seed reproduces
        ot = np.concatenate([other[r][s][0] for r in range(4)])
        assert not np.array_equal(np.sort(ot, axis=0), np.sort(all4, axis=0))            # another seed differs
        v = np.concatenate([four[r][s][1] for r in range(4)])
        n = v.size
        sig = cfg4.thermalVelocity[s]
        assert abs(v.mean()) < 5 * sig / np.sqrt(n)
        assert abs(v.std() / sig - 1) < 5 / np.sqrt(2 * n)
        assert abs(((v / sig) ** 4).mean() - 3) < 0.2                                    # Gaussian kurtosis
        # uniform positions: mean L/2, variance L^2/12 per dimension
        Lg = np.array(cfg4.nSubdomains) * np.array(cfg4.trueSize)
        assert np.all(np.abs(all4.mean(0) / Lg - 0.5) < 5 / np.sqrt(12 * len(all4)))


def test_device_start_steps_like_host_start(gpu_lib):
    """A cold-Langmuir run started on the device follows the host-started one (same particles up to ulps)."""
    from pinc_b200 import sim
    text, cfg = small_cfg("cold", grid__nsubdomains="1,1,1", grid__truesize="16,8,8", multigrid__mglevels=3,
                          population__nparticles="8 pc", population__nalloc="16 pc", population__perturbamplitude="2e-3,0,0,0,0,0")
    hist = []
    for mode in ("host", "device"):
        W = sim.World(cfg)
        if mode == "host":
            W.set_particles(initial.perturb(cfg, initial.lattice(cfg)))
        else:
            W.init_on_device("lattice")
        W.migrate(); W.field_solve(); W.half_kick()
        e = []
        for _ in range(4):
            W.step()
            e.append(W.energies())
        hist.append(np.array(e))
        W.close()
    assert np.abs(hist[0] - hist[1]).max() <= 1e-9 * np.abs(hist[0]).max()
