"""tcgen05 / TMEM plumbing self-test (GPU): a one-CTA bf16 GEMM through umma.cuh against torch fp32."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _selftest_lib():
    """tests/csrc/libtt_selftest.so - built by __graft_entry__.build(); test infrastructure, not in the product ABI."""
    import os
    from timbre_trap_b200 import build
    path = build.SELFTEST_OUT
    if not os.path.exists(path):
        build.build_selftest()
    lib = ctypes.CDLL(path)
    lib.tt_umma_probe.restype = ctypes.c_int
    lib.tt_umma_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    lib.tt_selftest_last_error.restype = ctypes.c_char_p
    lib.tt_umma_probe_mn.restype = ctypes.c_int
    lib.tt_umma_probe_mn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    return lib


def _run(n, k, swap):
    lib = _selftest_lib()
    g = torch.Generator(device='cuda').manual_seed(n * 1000 + k)
    a = torch.randn((128, k), device='cuda', generator=g).bfloat16()
    b = torch.randn((n, k), device='cuda', generator=g).bfloat16()
    d = torch.full((128, n), float('nan'), device='cuda')
    rc = lib.tt_umma_probe(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(d.data_ptr()), n, k, swap,
                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.tt_selftest_last_error()
    torch.cuda.synchronize()
    want = a.float() @ b.float().t()
    return float((d - want).abs().max()), float(want.abs().max())


@pytest.mark.parametrize('n,k', [(16, 16), (16, 32), (32, 64), (64, 128), (128, 256), (256, 64)])
def test_probe_gemm(n, k):
    err, scale = _run(n, k, 0)
    assert err <= 1e-3 * scale, (err, scale)


def _run_mn(n, k, order, shift=0, use=None):
    lib = _selftest_lib()
    use = k if use is None else use
    g = torch.Generator(device='cuda').manual_seed(n * 1000 + k + 7)
    a = torch.randn((128, k), device='cuda', generator=g).bfloat16()
    b = torch.randn((n, k), device='cuda', generator=g).bfloat16()
    d = torch.full((128, n), float('nan'), device='cuda')
    rc = lib.tt_umma_probe_mn(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(d.data_ptr()), n, k, order, shift, use,
                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.tt_selftest_last_error()
    torch.cuda.synchronize()
    want = a[:, shift:shift + use].float() @ b[:, :use].float().t()
    return float((d - want).abs().max()), float(want.abs().max())


@pytest.mark.parametrize('n,k', [(16, 16), (32, 64), (64, 128), (96, 256)])
def test_probe_gemm_mn_major(n, k):
    """MN-major operands ([group][k][8] - the C8 planar activation layout with pixels as the GEMM K axis), the form the tensor-core
    weight-gradient kernel uses: descriptor LBO = stride between 8-wide k groups, SBO = stride between 8-wide M / N groups."""
    err, scale = _run_mn(n, k, 0)
    assert err <= 1e-3 * scale, (err, scale)


def test_probe_gemm_mn_major_shifted_window():
    """A convolution tap = the A operand started a few 16-byte rows further along K."""
    for shift in (1, 3, 8, 13):
        err, scale = _run_mn(32, 160, 0, shift=shift, use=128)
        assert err <= 1e-3 * scale, (shift, err, scale)
