"""PINC .ini input for the stand-alone driver (tests, bench).

In a PINC build the ini layer stays the reference's host C (iniparser + src/io.c + src/units.c);
this module restates just enough of it for hosts that do not link that code:

* iniparser 3.1 semantics: `[section]`, `key = value`, `;`/`#` comments, keys lower-cased
  (/root/reference/lib/iniparser/src/iniparser.c:376), access as "section:key";
* list values split on ',' and cyclically expanded to the requested length
  (src/io.c:402-433 iniGetStrArr, :823 strArrExpand); numbers parsed with atof semantics
  (leading numeric prefix, so "64 pc" -> 64; src/io.c:332-408);
* suffixes `pc` (x global cell count) and `tot` (/ global length) (src/units.c:138-158,
  src/io.c:536-560);
* SI / semiSI normalisation so that dx = dt = eps0 = 1 (src/units.c:61-252).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

ELEMENTARY_CHARGE = 1.60217733e-19     # src/units.c:30
ELECTRON_MASS = 9.10938188e-31         # src/units.c:31
VACUUM_PERMITTIVITY = 8.854187817e-12  # src/units.c:32

_NUM = re.compile(r"^\s*[-+]?(0[xX][0-9a-fA-F.]+(?:[pP][-+]?\d+)?|(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)")


def atof(s: str) -> float:
    """C atof: numeric prefix, else 0.0 (also hex floats as written by iniSetDoubleArr)."""
    m = _NUM.match(s)
    if not m:
        return 0.0
    tok = m.group(0).strip()
    if "x" in tok.lower():
        return float.fromhex(tok)
    return float(tok)


class Ini:
    """Dictionary of "section:key" -> raw string, as iniparser_load builds it."""

    def __init__(self, text: str):
        self.d: dict[str, str] = {}
        sec = ""
        for raw in text.splitlines():
            line = raw.strip()
            if not line or line[0] in ";#":
                continue
            if line.startswith("[") and "]" in line:
                sec = line[1:line.index("]")].strip().lower()
                continue
            if "=" not in line:
                continue
            key, val = line.split("=", 1)
            for c in ";#":                       # trailing comment
                if c in val:
                    val = val[:val.index(c)]
            self.d[f"{sec}:{key.strip().lower()}"] = val.strip()

    @classmethod
    def load(cls, path: str, overrides: dict[str, str] | None = None) -> "Ini":
        with open(path) as f:
            ini = cls(f.read())
        for k, v in (overrides or {}).items():    # CLI "section:key=value" (src/io.c:273-276)
            ini.d[k.lower()] = str(v)
        return ini

    def dump(self) -> str:
        """ini text of the current dictionary (sections regrouped)."""
        secs: dict[str, list[str]] = {}
        for k, v in self.d.items():
            sec, key = k.split(":", 1)
            secs.setdefault(sec, []).append(f"{key} = {v}")
        return "\n".join(f"[{s}]\n" + "\n".join(lines) for s, lines in secs.items()) + "\n"

    def has(self, key: str) -> bool:
        return key.lower() in self.d

    def raw(self, key: str) -> str:
        key = key.lower()
        if key not in self.d:
            raise KeyError(f'Key "{key}" not found in input file')      # src/io.c:316-321
        return self.d[key]

    def n_elements(self, key: str) -> int:
        s = self.raw(key)
        return 0 if s == "" else s.count(",") + 1

    def str_arr(self, key: str, n: int) -> list[str]:
        items = [t.strip() for t in self.raw(key).split(",")]
        return [items[i % len(items)] for i in range(n)]

    def doubles(self, key: str, n: int) -> list[float]:
        return [atof(t) for t in self.str_arr(key, n)]

    def ints(self, key: str, n: int) -> list[int]:
        return [int(atof(t)) for t in self.str_arr(key, n)]

    def double(self, key: str) -> float:
        return atof(self.raw(key))

    def int(self, key: str) -> int:
        return int(atof(self.raw(key)))

    def set_doubles(self, key: str, vals) -> None:
        self.d[key.lower()] = ",".join(float(v).hex() for v in vals)

    def apply_suffix(self, key: str, suffix: str, mul: list[float]) -> None:
        if not self.has(key):
            return
        n = self.n_elements(key)
        out = []
        for i, t in enumerate(self.str_arr(key, n)):
            v = atof(t)
            if suffix in t:
                v *= mul[i % len(mul)]
            out.append(v)
        self.set_doubles(key, out)


@dataclass
class Config:
    nDims: int
    nSubdomains: list
    trueSize: list
    nGhostLayers: list
    thresholds: list
    boundaries: list
    nSpecies: int
    nParticles: list          # global, per species
    nAlloc: list              # global, per species
    charge: list              # normalised
    mass: list                # normalised
    thermalVelocity: list     # normalised (cells / step)
    drift: list
    perturbAmplitude: list    # normalised (cells), nSpecies*nDims
    perturbMode: list
    nEmigrantsAlloc: list
    BExt: list
    EExt: list
    nTimeSteps: int
    methods: dict = field(default_factory=dict)
    mgLevels: int = 1
    mgCycles: int = 1
    nPreSmooth: int = 10
    nPostSmooth: int = 10
    nCoarseSolve: int = 10
    units: dict = field(default_factory=dict)

    @property
    def nRanks(self) -> int:
        return math.prod(self.nSubdomains)

    @property
    def globalSize(self) -> list:
        return [a * b for a, b in zip(self.nSubdomains, self.trueSize)]


def _units_si(ini: Ini) -> dict:
    """src/units.c:191-231 (uSI) + :233-252 (uAddDerivedUnits)."""
    nDims = ini.int("grid:ndims")
    nS = ini.int("population:nspecies")
    T = ini.double("time:timestep")
    step = ini.doubles("grid:stepsize", nDims)
    nPart = [int(v) for v in ini.doubles("population:nparticles", nS)]
    dens = ini.doubles("population:density", nS)
    charge = ini.doubles("population:charge", nS)
    L = [a * b for a, b in zip(ini.ints("grid:nsubdomains", nDims), ini.ints("grid:truesize", nDims))]
    V = math.prod(L) * step[0] ** nDims
    weights = [dens[s] * V / nPart[s] for s in range(nS)]
    X = step[0]
    Q = weights[0] * abs(charge[0])
    M = (T * Q) ** 2 / (VACUUM_PERMITTIVITY * X ** nDims)
    u = dict(nDims=nDims, nSpecies=nS, weights=weights, length=X, time=T, charge=Q, mass=M)
    u["velocity"] = X / T
    u["density"] = 1.0 / X ** nDims
    u["eField"] = X * M / (T ** 2 * Q)
    u["bField"] = M / (T * Q)
    u["potential"] = (X / T) ** 2 * M / Q
    u["chargeDensity"] = Q / X ** nDims
    u["energy"] = M * (X / T) ** 2
    return u


def normalize(ini: Ini) -> dict:
    """uAlloc + uNormalize (src/units.c:61-136): rewrites the dictionary in place, returns units."""
    nDims = ini.int("grid:ndims")
    nS = ini.int("population:nspecies")
    L = [a * b for a, b in zip(ini.ints("grid:nsubdomains", nDims), ini.ints("grid:truesize", nDims))]
    V = float(math.prod(L))
    # parseIndirectInput (src/units.c:138-158)
    ini.apply_suffix("population:nparticles", "pc", [V])
    ini.apply_suffix("population:nalloc", "pc", [V])
    ini.apply_suffix("grid:nemigrantsalloc", "pc", [V])
    ini.apply_suffix("grid:stepsize", "tot", [1.0 / l for l in L])

    method = ini.raw("methods:normalization")
    if method == "semiSI":                                         # src/units.c:159-189
        charge = ini.doubles("population:charge", nS)
        mass = ini.doubles("population:mass", nS)
        dens = ini.doubles("population:density", nS)
        wpe = math.sqrt(ELEMENTARY_CHARGE ** 2 * dens[0] / (VACUUM_PERMITTIVITY * ELECTRON_MASS))
        ini.set_doubles("population:charge", [c * ELEMENTARY_CHARGE for c in charge])
        ini.set_doubles("population:mass", [m * ELECTRON_MASS for m in mass])
        ini.set_doubles("time:timestep", [ini.double("time:timestep") / wpe])
    elif method != "SI":
        raise ValueError("methods:normalization not valid (must be SI or semiSI)")
    u = _units_si(ini)

    w = u["weights"]
    # uNormalize multiplies by the reciprocal (adScale(.., 1.0/unit)); keep that rounding
    charge = [c0 * w[s] * (1.0 / u["charge"]) for s, c0 in enumerate(ini.doubles("population:charge", nS))]
    mass = [m0 * w[s] * (1.0 / u["mass"]) for s, m0 in enumerate(ini.doubles("population:mass", nS))]
    dens = [d0 / w[s] * (1.0 / u["density"]) for s, d0 in enumerate(ini.doubles("population:density", nS))]
    ini.set_doubles("population:charge", charge)
    ini.set_doubles("population:mass", mass)
    ini.set_doubles("population:density", dens)
    for key, unit in (("population:thermalvelocity", "velocity"), ("population:drift", "velocity"),
                      ("population:perturbamplitude", "length"), ("fields:bext", "bField"),
                      ("fields:eext", "eField")):
        if ini.has(key):
            n = ini.n_elements(key)
            ini.set_doubles(key, [v * (1.0 / u[unit]) for v in ini.doubles(key, n)])
    return u


def load_config(path_or_ini, overrides: dict | None = None) -> Config:
    ini = path_or_ini if isinstance(path_or_ini, Ini) else Ini.load(path_or_ini, overrides)
    u = normalize(ini)
    nDims = ini.int("grid:ndims")
    nS = ini.int("population:nspecies")

    def opt_doubles(key, n, default=0.0):
        return ini.doubles(key, n) if ini.has(key) else [default] * n

    methods = {k.split(":", 1)[1]: v for k, v in ini.d.items() if k.startswith("methods:")}
    return Config(
        nDims=nDims,
        nSubdomains=ini.ints("grid:nsubdomains", nDims),
        trueSize=ini.ints("grid:truesize", nDims),
        nGhostLayers=ini.ints("grid:nghostlayers", 2 * nDims),
        thresholds=ini.doubles("grid:thresholds", 2 * nDims),
        boundaries=ini.str_arr("grid:boundaries", 2 * nDims),
        nSpecies=nS,
        nParticles=[int(v) for v in ini.doubles("population:nparticles", nS)],
        nAlloc=[int(v) for v in ini.doubles("population:nalloc", nS)],
        charge=ini.doubles("population:charge", nS),
        mass=ini.doubles("population:mass", nS),
        # thermalVelocityCells: sigma_v given directly in cells/step (driver extension, not a PINC key)
        thermalVelocity=(ini.doubles("population:thermalvelocitycells", nS)
                         if ini.has("population:thermalvelocitycells")
                         else opt_doubles("population:thermalvelocity", nS)),
        drift=opt_doubles("population:drift", nS),
        perturbAmplitude=opt_doubles("population:perturbamplitude", nS * nDims),
        perturbMode=opt_doubles("population:perturbmode", nS * nDims),
        nEmigrantsAlloc=[int(v) for v in ini.doubles("grid:nemigrantsalloc", ini.n_elements("grid:nemigrantsalloc"))],
        BExt=opt_doubles("fields:bext", nDims),
        EExt=opt_doubles("fields:eext", nDims),
        nTimeSteps=ini.int("time:ntimesteps"),
        methods=methods,
        mgLevels=ini.int("multigrid:mglevels") if ini.has("multigrid:mglevels") else 1,
        mgCycles=ini.int("multigrid:mgcycles") if ini.has("multigrid:mgcycles") else 1,
        nPreSmooth=ini.int("multigrid:npresmooth") if ini.has("multigrid:npresmooth") else 10,
        nPostSmooth=ini.int("multigrid:npostsmooth") if ini.has("multigrid:npostsmooth") else 10,
        nCoarseSolve=ini.int("multigrid:ncoarsesolve") if ini.has("multigrid:ncoarsesolve") else 10,
        units=u,
    )
