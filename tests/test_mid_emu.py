"""CPU emulation of the fused 1024-point kernel (tests/emu/mid_emu.cpp): kernel_mid.cuh's lane functions, element maps
and twiddle pack -- the same source the GPU compiles -- run lane by lane on the host and must reproduce the oracle bit for
bit; the transposes must stay bank-conflict free with the changed phase-C map."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_SO = os.path.join(EMU_DIR, "liblwb_mid_emu.so")


def build_emu():
    src = os.path.join(EMU_DIR, "mid_emu.cpp")
    hdrs = [os.path.join(HERE, "..", "lewton_b200", "csrc", h) for h in ("kernel_mid.cuh", "kernel_long.cuh")]
    if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared",
                               "-o", EMU_SO, src])
    return EMU_SO


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(build_emu())
    vp = C.c_void_p
    L.lwb_emu_mid_build_pack.argtypes = [C.c_int] + [vp] * 5
    L.lwb_emu_mid_run.argtypes = [C.c_int, vp, vp, C.c_int, vp, vp, vp]
    return L


def P(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module", params=[1, 2], ids=["n1024", "n512"])
def kb_pack(request, emu, oracle):
    kb = request.param
    t = oracle.tables(11 - kb)
    pk = np.zeros(emu.lwb_emu_mid_pack_floats(), np.float32)
    emu.lwb_emu_mid_build_pack(kb, P(t.a), P(t.b), P(t.c), P(t.window), P(pk))
    return kb, pk


def oracle_run(oracle, bs, spec, state):
    pwr = oracle.Pwr(1, bs)
    if state is not None:
        pwr.set_data(state[None, :])
    outs = []
    for p in range(spec.shape[0]):
        rc, pcm = oracle.synth_spectrum(bs, bs, 1, 1, 1, spec[p:p + 1], pwr)
        assert rc == 0
        outs.append(pcm[0])
    return np.concatenate(outs), pwr.data()[0]


@pytest.mark.parametrize("seed,npk,prev,scale", [(20, 4, (0, 0, 1, 0), 1.0), (21, 3, (1, 0, 0, 1), 1.0), (22, 1, (0, 1, 1, 1), 1e-2),
                                                 (23, 5, (1, 1, 0, 0), 1.0), (24, 2, (1, 1, 1, 1), 1e-30), (25, 2, (0, 0, 0, 0), 1e30)])
def test_emulated_mid_kernel_matches_oracle(emu, kb_pack, oracle, seed, npk, prev, scale):
    """2^KB runs in lockstep, like a warp of k_mid; imported states are NOT symmetric."""
    kb, pack = kb_pack
    nb, n2, bs = 1 << kb, 1024 >> kb, 11 - kb
    rng = np.random.default_rng(seed)
    spec = (rng.standard_normal((nb, npk, n2)) * scale).astype(np.float32)
    states = (rng.standard_normal((nb, n2)) * scale).astype(np.float32)
    st = states.copy()
    out = np.zeros((nb, npk, n2), np.float32)
    hp = np.array(prev[:nb], np.int32)
    conflicts = emu.lwb_emu_mid_run(kb, P(pack), P(spec), npk, P(hp), P(st), P(out))
    assert conflicts == 1, "shared-memory transposes must be bank-conflict free"
    for b in range(nb):
        want, want_state = oracle_run(oracle, bs, spec[b], states[b] if hp[b] else None)
        emitted = npk if hp[b] else npk - 1
        got = out[b, :emitted].ravel()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), b
        assert np.array_equal(st[b].view(np.uint32), want_state.view(np.uint32)), b


def test_emulated_mid_kernel_special_values(emu, kb_pack, oracle):
    kb, pack = kb_pack
    nb, n2, bs = 1 << kb, 1024 >> kb, 11 - kb
    rng = np.random.default_rng(9)
    spec = rng.standard_normal((nb, 3, n2)).astype(np.float32)
    spec[0, 0, :64] = 1e-42          # denormals
    spec[0, 1, 5] = np.inf
    spec[1, 1, 77] = np.nan
    spec[1, 2] = 0.0
    st = np.zeros((nb, n2), np.float32)
    out = np.zeros((nb, 3, n2), np.float32)
    emu.lwb_emu_mid_run(kb, P(pack), P(spec), 3, P(np.zeros(nb, np.int32)), P(st), P(out))
    for b in range(nb):
        want, _ = oracle_run(oracle, bs, spec[b], None)
        got = out[b, :2].ravel()
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert np.all(same), b
