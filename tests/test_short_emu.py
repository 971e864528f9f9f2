"""CPU emulation of the fused short-block kernel (tests/emu/short_emu.cpp): kernel_short.cuh's per-lane phase
functions, element maps, swizzle, twiddle pack and neighbour-lane overlap hand-over -- the same source the GPU
compiles -- run lane by lane on the host and must reproduce the oracle bit for bit (imdct.rs:291-659 for
n = 256, audio.rs:1079-1154)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_SO = os.path.join(EMU_DIR, "liblwb_short_emu.so")


def build_emu():
    src = os.path.join(EMU_DIR, "short_emu.cpp")
    hdrs = [os.path.join(HERE, "..", "lewton_b200", "csrc", f) for f in ("kernel_short.cuh", "kernel_long.cuh")]
    if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared",
                               "-o", EMU_SO, src])
    return EMU_SO


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(build_emu())
    vp = C.c_void_p
    L.lwb_emu_short_build_pack.argtypes = [vp] * 5
    L.lwb_emu_short_run.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    return L


def P(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def pack(emu, oracle):
    t = oracle.tables(8)
    pk = np.zeros(emu.lwb_emu_short_pack_floats(), np.float32)
    emu.lwb_emu_short_build_pack(P(t.a), P(t.b), P(t.c), P(t.window), P(pk))
    return pk


def oracle_run(oracle, spec, state):
    """Short blocks of a bs0 = 8 / bs1 = 11 stream, one channel."""
    pwr = oracle.Pwr(1, 11)
    if state is not None:
        pwr.set_data(state[None, :])
    outs = []
    for p in range(spec.shape[0]):
        rc, pcm = oracle.synth_spectrum(8, 11, 0, 1, 1, spec[p:p + 1], pwr)
        assert rc == 0
        outs.append(pcm[0])
    return np.concatenate(outs), pwr.data()[0]


@pytest.mark.parametrize("seed,npk,with_state,scale", [(0, 8, False, 1.0), (1, 1, False, 1.0), (2, 1, True, 1.0), (3, 5, True, 1e-2),
                                                       (4, 9, True, 1.0), (5, 16, False, 1e-30), (6, 23, True, 1e30),
                                                       (7, 3, False, 1.0), (8, 64, True, 1.0)])
def test_emulated_short_kernel_matches_oracle(emu, pack, oracle, seed, npk, with_state, scale):
    rng = np.random.default_rng(seed)
    spec = (rng.standard_normal((npk, 128)) * scale).astype(np.float32)
    state = (rng.standard_normal(128) * scale).astype(np.float32) if with_state else None      # NOT symmetric
    want, want_state = oracle_run(oracle, spec, state)
    st = state.copy() if with_state else np.zeros(128, np.float32)
    out = np.zeros((npk, 128), np.float32)
    conflicts = emu.lwb_emu_short_run(P(pack), P(spec), npk, int(with_state), P(st), P(out))
    assert conflicts == 1, "tile reads and the transpose must be bank-conflict free"
    emitted = npk if with_state else npk - 1
    got = out[:emitted].ravel()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.nonzero(got.view(np.uint32) != want.view(np.uint32))[0][:8]
    assert np.array_equal(st.view(np.uint32), want_state.view(np.uint32))


def test_emulated_short_kernel_special_values(emu, pack, oracle):
    rng = np.random.default_rng(19)
    spec = rng.standard_normal((11, 128)).astype(np.float32)
    spec[0, :16] = 1e-42          # denormals are kept
    spec[3, 5] = np.inf
    spec[4, 77] = np.nan
    spec[9] = 0.0
    want, want_state = oracle_run(oracle, spec, None)
    st = np.zeros(128, np.float32)
    out = np.zeros((11, 128), np.float32)
    emu.lwb_emu_short_run(P(pack), P(spec), 11, 0, P(st), P(out))
    got = out[:10].ravel()
    assert np.all((got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want)))


@pytest.mark.parametrize("esz", [4, 2])
def test_short_kernel_pcm_staging_is_consistent_and_conflict_free(emu, esz):
    """The swizzled PCM staging of k_short: every staged sample sits where the per-packet vector copy reads it,
    and neither the 32 staging stores nor the vector loads have shared-memory bank conflicts."""
    assert emu.lwb_emu_short_staging_conflicts(esz) == 1
