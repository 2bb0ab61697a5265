"""Initial conditions for the stand-alone driver (host side, runs once at t=0).

Restates pPosLattice / pPosPerturb / pVelZero (src/population.c:172-276, 412-428) and provides a
seeded uniform + Maxwellian start (pPosUniform :110-170, pVelMaxwell :367-392) with numpy's Philox
generator instead of GSL's MT19937 + ziggurat, which cannot be reproduced without GSL
(SURVEY 8d "Seeds/distributions").  Positions are returned per rank in the LOCAL frame
(x_local = x_global - offset, offset = subdomain*trueSize - nGhost; src/grid.c:525).
"""
from __future__ import annotations

import math

import numpy as np


def rank_subdomain(rank, nSub):
    """src/grid.c:166-171: x fastest."""
    sub = []
    for d in range(len(nSub)):
        sub.append(rank % nSub[d])
        rank //= nSub[d]
    return sub


def rank_offset(rank, cfg):
    sub = rank_subdomain(rank, cfg.nSubdomains)
    return [sub[d] * cfg.trueSize[d] - cfg.nGhostLayers[d] for d in range(cfg.nDims)]


def _owner_split(glob_pos, vel, cfg):
    """Distribute global-frame particles to ranks: owner = floor(x / trueSize) per dimension
    (posToSubdomain = 1/trueSize, src/grid.c:526 and population.c:143-146)."""
    nD = cfg.nDims
    own = np.zeros(len(glob_pos), dtype=np.int64)
    mul = 1
    for d in range(nD):
        sd = (glob_pos[:, d] * (1.0 / cfg.trueSize[d])).astype(np.int32)
        own += sd * mul
        mul *= cfg.nSubdomains[d]
    out = []
    for r in range(cfg.nRanks):
        m = own == r
        off = np.array(rank_offset(r, cfg), dtype=np.float64)
        out.append((glob_pos[m] - off, vel[m]))
    return out


def lattice(cfg):
    """pPosLattice + pVelZero: per rank, list over species of (pos_local[n,3], vel[n,3])."""
    L = cfg.globalSize
    V = math.prod(L)
    nD = cfg.nDims
    per_rank = [[] for _ in range(cfg.nRanks)]
    for s in range(cfg.nSpecies):
        n = cfg.nParticles[s]
        l = (V / float(n)) ** (1.0 / nD)
        lin = l * np.arange(n, dtype=np.float64)
        pos = np.empty((n, nD))
        for d in range(nD):
            pos[:, d] = np.fmod(lin, L[d])
            lin = lin / L[d]
        split = _owner_split(pos, np.zeros_like(pos), cfg)
        for r in range(cfg.nRanks):
            per_rank[r].append(split[r])
    return per_rank


def perturb(cfg, per_rank):
    """pPosPerturb: x += A cos(2 pi m x / L) in the global frame, per species and dimension."""
    L = cfg.globalSize
    nD = cfg.nDims
    for r in range(cfg.nRanks):
        off = np.array(rank_offset(r, cfg), dtype=np.float64)
        for s in range(cfg.nSpecies):
            pos, _ = per_rank[r][s]
            pos += off                                       # pToGlobalFrame
            for d in range(nD):
                theta = 2.0 * math.pi * cfg.perturbMode[s * nD + d] * pos[:, d] / L[d]
                pos[:, d] += cfg.perturbAmplitude[s * nD + d] * np.cos(theta)
            pos -= off                                       # pToLocalFrame
    return per_rank


def maxwellian(cfg, seed=20261018, n_particles=None, ranks=None):
    """Uniform positions + Maxwellian velocities, seeded; |v| components redrawn until < maxVel=1.

    Every rank draws only its own particles (Philox keyed by (seed, rank)), uniformly inside its own
    sub-domain, so set-up cost does not grow with the number of ranks.  n_particles: global count
    per species (default cfg.nParticles); each rank gets n/nRanks.
    """
    nD = cfg.nDims
    n_particles = n_particles or cfg.nParticles
    out = []
    for r in (range(cfg.nRanks) if ranks is None else ranks):
        rng = np.random.Generator(np.random.Philox(key=[seed, r]))
        species = []
        for s in range(cfg.nSpecies):
            n = n_particles[s] // cfg.nRanks
            pos = np.empty((n, nD))
            for d in range(nD):
                # local frame: true cells span [nGhost, nGhost + trueSize)
                pos[:, d] = cfg.nGhostLayers[d] + cfg.trueSize[d] * rng.random(n)
            vel = cfg.thermalVelocity[s] * rng.standard_normal((n, nD)) + cfg.drift[s]
            bad = np.abs(vel) >= 1.0
            while bad.any():
                vel[bad] = cfg.thermalVelocity[s] * rng.standard_normal(int(bad.sum())) + cfg.drift[s]
                bad = np.abs(vel) >= 1.0
            species.append((pos, vel))
        out.append(species)
    return out
