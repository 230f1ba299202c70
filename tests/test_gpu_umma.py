"""GPU tests of the tcgen05/TMEM/TMA bf16 GEMM (avd_gemm_bf16) against a plain PyTorch fp32 reference of the
same op on the same bf16-rounded operands (so the only difference is accumulation order: tol 1e-3 normwise,
typically ~1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from avddpg_b200 import _lib
    _lib.require_device()
    return _lib


def _run(lib, layout, batch, M, N, K, splitk=1, lda=None, ldb=None):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + layout)
    if layout == 0:
        lda, ldb = lda or K, ldb or K
        A = torch.randn(batch, M, lda, device="cuda", generator=g).bfloat16()
        B = torch.randn(batch, N, ldb, device="cuda", generator=g).bfloat16()
        ref = torch.einsum("bmk,bnk->bmn", A[..., :K].float(), B[..., :K].float())
        a_batch, b_batch = M * lda, N * ldb
    else:
        lda, ldb = lda or M, ldb or N
        A = torch.randn(batch, K, lda, device="cuda", generator=g).bfloat16()
        B = torch.randn(batch, K, ldb, device="cuda", generator=g).bfloat16()
        ref = torch.einsum("bkm,bkn->bmn", A[..., :M].float(), B[..., :N].float())
        a_batch, b_batch = K * lda, K * ldb
    C = torch.zeros(batch, M, N, device="cuda")
    lib.check(lib.load().avd_gemm_bf16(layout, batch, M, N, K, lib.ptr(A), lda, a_batch, lib.ptr(B), ldb, b_batch, lib.ptr(C), N, M * N,
                                       splitk, lib.current_stream()))
    torch.cuda.synchronize()
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    return err


@pytest.mark.parametrize("batch,M,N,K", [(1, 128, 128, 64), (1, 128, 128, 256), (2, 256, 128, 320), (1, 200, 128, 304), (3, 70, 128, 256),
                                         (1, 1000, 304, 128), (2, 128, 48, 128)])
def test_tn_forward_and_dgrad_shapes(lib, batch, M, N, K):
    assert _run(lib, 0, batch, M, N, K) < 1e-3


@pytest.mark.parametrize("batch,M,N,K,splitk", [(1, 128, 128, 64, 1), (1, 256, 128, 512, 1), (2, 304, 128, 1000, 4), (1, 256, 128, 4096, 8),
                                                (4, 304, 128, 64, 1)])
def test_nt_wgrad_shapes(lib, batch, M, N, K, splitk):
    assert _run(lib, 1, batch, M, N, K, splitk) < 1e-3


def test_tn_splitk_accumulates(lib):
    assert _run(lib, 0, 1, 128, 128, 1024, splitk=4) < 1e-3


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("fmt", [0, 1])
def test_fp16_and_bf16_operands(lib, layout, fmt):
    """kind::f16 with fp16 (0) or bf16 (1) operands: the product of the rounded operands.  precision = 2 of the learn step uses fp16."""
    batch, M, N, K = 2, 256, 128, 320
    g = torch.Generator(device="cuda").manual_seed(17 + fmt + 4 * layout)
    dt = {0: torch.float16, 1: torch.bfloat16}[fmt]
    if layout == 0:
        A = torch.randn(batch, M, K, device="cuda", generator=g).to(dt)
        B = torch.randn(batch, N, K, device="cuda", generator=g).to(dt)
        ref = torch.einsum("bmk,bnk->bmn", A.float(), B.float())
        lda, ldb, a_batch, b_batch = K, K, M * K, N * K
    else:
        A = torch.randn(batch, K, M, device="cuda", generator=g).to(dt)
        B = torch.randn(batch, K, N, device="cuda", generator=g).to(dt)
        ref = torch.einsum("bkm,bkn->bmn", A.float(), B.float())
        lda, ldb, a_batch, b_batch = M, N, K * M, K * N
    C = torch.zeros(batch, M, N, device="cuda")
    lib.check(lib.load().avd_gemm_f16kind(layout, batch, M, N, K, lib.ptr(A), lda, a_batch, lib.ptr(B), ldb, b_batch, lib.ptr(C), N, M * N, 1,
                                          fmt, fmt, lib.current_stream()))
    torch.cuda.synchronize()
    assert (C - ref).abs().max().item() / ref.abs().max().item() < 1e-3


def test_mixed_operand_formats_are_refused(lib):
    """fp16 x bf16 is expressible in the instruction descriptor but traps on B200 (measured: cudaErrorIllegalInstruction), so the
    entry point refuses it instead of poisoning the context."""
    A = torch.zeros(128, 64, device="cuda", dtype=torch.float16)
    B = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    C = torch.zeros(128, 128, device="cuda")
    with pytest.raises(NotImplementedError):
        lib.check(lib.load().avd_gemm_f16kind(0, 1, 128, 128, 64, lib.ptr(A), 64, 128 * 64, lib.ptr(B), 64, 128 * 64, lib.ptr(C), 128, 128 * 128, 1,
                                              0, 1, lib.current_stream()))
