"""ctypes binding of pinc_b200/libpinc_b200.so (the C-ABI of include/pinc_b200.h).

This is the only way the Python host code reaches the numerics: there is no Python/CPU fallback.  If the
shared library is missing, loading raises; if it is present but no CUDA device is usable, the first entry
point that needs the device prints "PINC-B200 ERROR: ..." and exits, like msg(ERROR, ...) in the reference
(/root/reference/src/io.c:170-217).
"""
from __future__ import annotations

import ctypes as C
import os

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PINC_B200_LIB") or os.path.join(HERE, "libpinc_b200.so")      # the override is for A/B builds of the same ABI

P = C.POINTER
_lib = None

# name -> (restype, argtypes); every symbol include/pinc_b200.h declares
SIGNATURES = {
    # particle path
    "puMove": (None, [P(abi.Population), C.c_void_p]),
    "puAcc3D1": (None, [P(abi.Population), P(abi.Grid)]),
    "puAcc3D1KE": (None, [P(abi.Population), P(abi.Grid)]),
    "puBoris3D1": (None, [P(abi.Population), P(abi.Grid), abi.c_double_p, abi.c_double_p]),
    "puBoris3D1KE": (None, [P(abi.Population), P(abi.Grid), abi.c_double_p, abi.c_double_p]),
    "puDistr3D1": (None, [P(abi.Population), P(abi.Grid)]),
    "puExtractEmigrants3D": (None, [P(abi.Population), P(abi.MpiInfo)]),
    "puMigrate": (None, [P(abi.Population), P(abi.MpiInfo), P(abi.Grid)]),
    "puRankToNeighbor": (C.c_int, [P(abi.MpiInfo), C.c_int]),
    "puNeighborToRank": (C.c_int, [P(abi.MpiInfo), C.c_int]),
    "puNeighborToReciprocal": (C.c_int, [C.c_int, C.c_int]),
    "pSumKinEnergy": (None, [P(abi.Population)]),
    "pPosAssertInLocalFrame": (None, [P(abi.Population), P(abi.Grid)]),
    "pVelAssertMax": (None, [P(abi.Population), C.c_double]),
    "pNew": (None, [P(abi.Population), C.c_int, abi.c_double_p, abi.c_double_p]),
    "pCut": (None, [P(abi.Population), C.c_int, C.c_long, abi.c_double_p, abi.c_double_p]),
    "pincGet3DRotationParameters": (None, [C.c_int, abi.c_double_p, abi.c_double_p, abi.c_double_p, abi.c_double_p, abi.c_double_p]),
    "pincPuSanity": (C.c_int, [C.c_char_p, C.c_int, abi.c_int_p, abi.c_double_p, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    # grid path
    "getSlice": (None, [abi.c_double_p, P(abi.Grid), C.c_int, C.c_int]),
    "setSlice": (None, [abi.c_double_p, P(abi.Grid), C.c_int, C.c_int]),
    "addSlice": (None, [abi.c_double_p, P(abi.Grid), C.c_int, C.c_int]),
    "gHaloOp": (None, [C.c_void_p, P(abi.Grid), P(abi.MpiInfo), C.c_int]),
    "gHaloOpDim": (None, [C.c_void_p, P(abi.Grid), P(abi.MpiInfo), C.c_int, C.c_int]),
    "gFinDiff1st": (None, [P(abi.Grid), P(abi.Grid)]),
    "gFinDiff2nd3D": (None, [P(abi.Grid), P(abi.Grid)]),
    "gZero": (None, [P(abi.Grid)]),
    "gMul": (None, [P(abi.Grid), C.c_double]),
    "gAdd": (None, [P(abi.Grid), C.c_double]),
    "gSub": (None, [P(abi.Grid), C.c_double]),
    "gSquare": (None, [P(abi.Grid)]),
    "gCopy": (None, [P(abi.Grid), P(abi.Grid)]),
    "gAddTo": (None, [P(abi.Grid), P(abi.Grid)]),
    "gSubFrom": (None, [P(abi.Grid), P(abi.Grid)]),
    "gSumTruegrid": (C.c_double, [P(abi.Grid)]),
    "gTotTruesize": (C.c_long, [P(abi.Grid), P(abi.MpiInfo)]),
    "gNeutralizeGrid": (None, [P(abi.Grid), P(abi.MpiInfo)]),
    "gBnd": (None, [P(abi.Grid), P(abi.MpiInfo)]),
    "gDirichlet": (None, [P(abi.Grid), C.c_int, P(abi.MpiInfo)]),
    "gNeumann": (None, [P(abi.Grid), C.c_int, P(abi.MpiInfo)]),
    "gSetBndSlices": (None, [P(abi.Grid), P(abi.MpiInfo)]),
    "mgRestrictBnd": (None, [P(abi.Multigrid)]),
    "gPotEnergy": (None, [P(abi.Grid), P(abi.Grid), P(abi.Population)]),
    # multigrid
    "mgSolve": (None, [P(abi.MultigridSolver), P(abi.Grid), P(abi.Grid), P(abi.MpiInfo)]),
    "mgSolveRaw": (None, [C.c_void_p, P(abi.Multigrid), P(abi.Multigrid), P(abi.Multigrid), P(abi.MpiInfo)]),
    "mgVRecursive": (None, [C.c_int, C.c_int, C.c_int, P(abi.Multigrid), P(abi.Multigrid), P(abi.Multigrid), P(abi.MpiInfo)]),
    "mgGS3D": (None, [P(abi.Grid), P(abi.Grid), C.c_int, P(abi.MpiInfo)]),
    "mgJacob3D": (None, [P(abi.Grid), P(abi.Grid), C.c_int, P(abi.MpiInfo)]),
    "mgVRegular": (None, [C.c_int, C.c_int, C.c_int, P(abi.Multigrid), P(abi.Multigrid), P(abi.Multigrid), P(abi.MpiInfo)]),
    "mgW": (None, [C.c_int, C.c_int, C.c_int, P(abi.Multigrid), P(abi.Multigrid), P(abi.Multigrid), P(abi.MpiInfo)]),
    "mgFMG": (None, [C.c_int, C.c_int, C.c_int, P(abi.Multigrid), P(abi.Multigrid), P(abi.Multigrid), P(abi.MpiInfo)]),
    "mgHalfRestrict3D": (None, [P(abi.Grid), P(abi.Grid)]),
    "mgBilinProl3D": (None, [P(abi.Grid), P(abi.Grid), P(abi.MpiInfo)]),
    "mgResidual": (None, [P(abi.Grid), P(abi.Grid), P(abi.Grid), P(abi.MpiInfo)]),
    "mgSumTrueSquared": (C.c_double, [P(abi.Grid), P(abi.MpiInfo)]),
    "mgSolver": (None, [P(C.c_void_p), P(C.c_void_p), P(C.c_void_p)]),
    "pincMgAllocSolver": (P(abi.MultigridSolver), [P(abi.Grid), P(abi.Grid), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mgFreeSolver": (None, [P(abi.MultigridSolver)]),
    "pincMgLastHistory": (C.c_int, [abi.c_double_p, C.c_int]),
    "pincMgSetMode": (None, [C.c_int]),
    "pincMgSetReplica": (None, [C.c_int]),
    "pincMgSetRowMode": (None, [C.c_int]),
    "puAccND1": (None, [P(abi.Population), P(abi.Grid)]),
    "puAccND1KE": (None, [P(abi.Population), P(abi.Grid)]),
    "puAccND0": (None, [P(abi.Population), P(abi.Grid)]),
    "puAccND0KE": (None, [P(abi.Population), P(abi.Grid)]),
    "puDistrND1": (None, [P(abi.Population), P(abi.Grid)]),
    "puDistrND0": (None, [P(abi.Population), P(abi.Grid)]),
    "puExtractEmigrantsND": (None, [P(abi.Population), P(abi.MpiInfo)]),
    "pincSetSlotted": (None, [C.c_int, C.c_int, C.c_int]),
    "pincPopLayout": (C.c_int, [P(abi.Population)]),
    "pincSlottedOverflows": (C.c_long, []),
    "pincMgSetHybrid": (None, [C.c_int]),
    "pincMgLastBarRes": (C.c_double, []),
    "pincMgLastPath": (C.c_int, []),
    # entry points that take PINC's dictionary *ini (need the host's iniGet*; typed for the symbol check)
    "puAcc3D1_set": (C.c_void_p, [C.c_void_p]),
    "puAcc3D1KE_set": (C.c_void_p, [C.c_void_p]),
    "puDistr3D1_set": (C.c_void_p, [C.c_void_p]),
    "puAccND1_set": (C.c_void_p, [C.c_void_p]),
    "puAccND1KE_set": (C.c_void_p, [C.c_void_p]),
    "puAccND0_set": (C.c_void_p, [C.c_void_p]),
    "puAccND0KE_set": (C.c_void_p, [C.c_void_p]),
    "puDistrND1_set": (C.c_void_p, [C.c_void_p]),
    "puDistrND0_set": (C.c_void_p, [C.c_void_p]),
    "puExtractEmigrantsND_set": (C.c_void_p, [C.c_void_p]),
    "puExtractEmigrants3D_set": (C.c_void_p, [C.c_void_p]),
    "puGet3DRotationParameters": (None, [C.c_void_p, abi.c_double_p, abi.c_double_p]),
    "mgSolver_set": (C.c_void_p, [C.c_void_p]),
    "mgAllocSolver": (P(abi.MultigridSolver), [C.c_void_p, P(abi.Grid), P(abi.Grid)]),
    # host-struct constructors
    "pincGridAlloc": (P(abi.Grid), [C.c_int, abi.c_int_p, abi.c_int_p, C.c_int, abi.c_int_p]),
    "pincGridFree": (None, [P(abi.Grid)]),
    "pincMpiAlloc": (P(abi.MpiInfo), [C.c_int, C.c_int, abi.c_int_p, abi.c_int_p, abi.c_int_p, C.c_int, C.c_int]),
    "pincMpiFree": (None, [P(abi.MpiInfo)]),
    "pincCreateNeighborhood": (None, [P(abi.MpiInfo), P(abi.Grid), abi.c_long_p, C.c_int, abi.c_double_p]),
    "pincPopAlloc": (P(abi.Population), [C.c_int, C.c_int, abi.c_long_p, abi.c_double_p, abi.c_double_p]),
    "pincPopFree": (None, [P(abi.Population)]),
    # initial conditions on the device
    "pincPosLattice": (None, [P(abi.Population), P(abi.MpiInfo), abi.c_long_p, abi.c_int_p]),
    "pincPosUniform": (None, [P(abi.Population), P(abi.MpiInfo), abi.c_long_p, abi.c_int_p, C.c_ulonglong]),
    "pincPosPerturb": (None, [P(abi.Population), P(abi.MpiInfo), abi.c_double_p, abi.c_double_p, abi.c_int_p]),
    "pincVelMaxwell": (None, [P(abi.Population), P(abi.MpiInfo), abi.c_double_p, abi.c_double_p, C.c_ulonglong]),
    "pincVelZero": (None, [P(abi.Population)]),
    # context, coherence, transport, timing
    "pincCtxCreate": (C.c_void_p, [C.c_int, C.c_int, C.c_int]),
    "pincCtxMakeCurrent": (None, [C.c_void_p]),
    "pincCtxDestroy": (None, [C.c_void_p]),
    "pincCommInitThreads": (None, [P(C.c_void_p), C.c_int]),
    "pincNcclUniqueId": (C.c_int, [C.c_char_p]),
    "pincCommInitNccl": (None, [C.c_void_p, C.c_char_p]),
    "pincTransportName": (C.c_char_p, []),
    "pincSyncGridToDevice": (None, [P(abi.Grid)]),
    "pincSyncGridToHost": (None, [P(abi.Grid)]),
    "pincSyncPopToDevice": (None, [P(abi.Population)]),
    "pincSyncPopToHost": (None, [P(abi.Population)]),
    "pincForget": (None, [C.c_void_p]),
    "pincHostRegister": (C.c_int, [C.c_void_p, C.c_size_t]),
    "pincHostUnregister": (C.c_int, [C.c_void_p]),
    "pincDeviceSynchronize": (None, []),
    "pincAccMove3D1KE": (None, [P(abi.Population), P(abi.Grid), P(abi.MpiInfo)]),
    "pincAccMoveDistr3D1KE": (None, [P(abi.Population), P(abi.Grid), P(abi.Grid), P(abi.MpiInfo)]),
    "pincTimerStart": (None, []),
    "pincTimerStopMs": (C.c_double, []),
    "pincProfEnable": (None, [C.c_int]),
    "pincProfReset": (None, []),
    "pincProfGet": (C.c_int, [C.c_int, C.c_char_p, C.c_int, abi.c_double_p, abi.c_long_p, abi.c_double_p]),
    "pincLaunchCount": (C.c_long, []),
    "pincVersion": (C.c_char_p, []),
    "pincLastError": (C.c_int, [C.c_char_p, C.c_int]),
}


def load():
    """Load libpinc_b200.so and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C pinc_b200/csrc).  pinc_b200 has no CPU fallback.")
    lib = C.CDLL(SO_PATH, mode=os.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("PINC_B200_LIB") and not hasattr(lib, name):
            continue                     # an older A/B build of the library
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def fn_ptr(lib, name) -> C.c_void_p:
    return C.cast(getattr(lib, name), C.c_void_p)


def profile(lib) -> dict:
    """Per-kernel-class accounting since the last pincProfReset: {class: (ms, launches, algorithmic bytes)}."""
    out = {}
    name = C.create_string_buffer(32)
    ms, cnt, by = C.c_double(), C.c_long(), C.c_double()
    i = 0
    while lib.pincProfGet(i, name, 32, C.byref(ms), C.byref(cnt), C.byref(by)):
        if cnt.value:
            out[name.value.decode()] = (ms.value, cnt.value, by.value)
        i += 1
    return out
