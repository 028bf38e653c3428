"""tcgen05 shared-memory descriptor probe (diagnostic, GPU): which address/swizzle rule does the hardware apply
when an operand is an overlapping ("Toeplitz") view of a linear signal?  Drives zns_dbg_umma_raw with hand-built
shared-memory images and compares the accumulator with numpy predictions under two hypotheses:
  abs: the 16-byte-chunk XOR is a function of the final absolute address bits (bits 7..9 -> bits 4..6 for SW128)
  rel: the XOR uses the row index inside the 8-row group (+ the descriptor's base_offset field)
Also times MMA sequences for several N / layouts (cycles per 128 x N x 16 MMA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from zeronotesamba_b200 import _lib as L

A_BYTES = 96 * 1024      # operand-A region of the image
B_OFF = A_BYTES
IMG = 160 * 1024

SW = {"none": (0, 16, 0), "sw32": (6, 32, 1), "sw64": (4, 64, 3), "sw128": (2, 128, 7)}  # layout code, pitch, xor mask


def desc(start, lbo, sbo, layout, base_offset=0):
    d = (start >> 4) & 0x3FFF
    d |= ((lbo >> 4) & 0x3FFF) << 16
    d |= ((sbo >> 4) & 0x3FFF) << 32
    d |= 1 << 46
    d |= (base_offset & 7) << 49
    d |= layout << 61
    return d


def idesc_f16(m, n):
    return (1 << 4) | ((n >> 3) << 17) | ((m >> 4) << 24)


def put_b(img, Bm):
    """B [N][K] fp16 -> canonical K-major SW128 tiles at B_OFF; returns per-k-step start offsets."""
    N, K = Bm.shape
    h = img.view(np.float16)
    for n in range(N):
        for k in range(K):
            kb, kl = divmod(k, 64)
            off = B_OFF + kb * (N * 128) + (n // 8) * 1024 + (n % 8) * 128 + (((kl // 8) ^ (n % 8)) * 16) + (kl % 8) * 2
            h[off // 2] = Bm[n, k]
    return [B_OFF + (ks // 4) * (N * 128) + (ks % 4) * 32 for ks in range(K // 16)]


def run(img, mmas, n_cols, reps=1):
    """mmas: list of (a_desc, b_desc, col, acc, idesc)."""
    dev = "cuda"
    t_img = torch.from_numpy(img.copy()).to(dev)
    a = torch.tensor([m[0] for m in mmas], dtype=torch.int64).to(dev)   # bit patterns < 2^63 except layout bits
    b = torch.tensor([m[1] for m in mmas], dtype=torch.int64).to(dev)
    col = torch.tensor([m[2] for m in mmas], dtype=torch.int32).to(dev)
    acc = torch.tensor([m[3] for m in mmas], dtype=torch.int32).to(dev)
    ids = torch.tensor([m[4] for m in mmas], dtype=torch.int64).to(torch.int32).to(dev)
    out = torch.zeros(128, n_cols, dtype=torch.float32, device=dev)
    cyc = torch.zeros(1, dtype=torch.int64, device=dev)
    L.check(L.lib().zns_dbg_umma_raw(L.ptr(t_img), img.nbytes, L.ptr(a), L.ptr(b), L.ptr(col), L.ptr(acc), L.ptr(ids),
                                     len(mmas), n_cols, L.ptr(out), reps, L.ptr(cyc), L.current_stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy(), int(cyc.item())


def s64(x):
    """python int bit pattern -> signed int64 value torch accepts"""
    return x - (1 << 64) if x >= (1 << 63) else x


def toeplitz_test(mode, hop_elems, start_elems, K, N, base_offset_rule, rng):
    """A(r, k) = S[hop*r + start + k] as a view of a linear signal stored with the absolute-address swizzle."""
    layout, pitch, mask = SW[mode]
    assert hop_elems * 2 == pitch
    S = rng.integers(-3, 4, size=A_BYTES // 2).astype(np.float16)
    img = np.zeros(IMG, dtype=np.uint8)
    h = img.view(np.float16)
    for p in range(S.size):
        logical = 2 * p
        phys = logical ^ ((((logical >> 7) & mask)) << 4)
        h[phys // 2] = S[p]
    Bm = rng.integers(-3, 4, size=(N, K)).astype(np.float16)
    b_starts = put_b(img, Bm)
    mmas = []
    for ks in range(K // 16):
        st = 2 * (start_elems + 16 * ks)
        bo = ((st >> 7) & 7) if base_offset_rule else 0
        mmas.append((s64(desc(st, 16, 8 * pitch, layout, bo)), s64(desc(b_starts[ks], 16, 1024, 2)), 0, int(ks > 0),
                     idesc_f16(128, N)))
    D, _ = run(img, mmas, 32 * ((N + 31) // 32))
    D = D[:, :N]
    r = np.arange(128)[:, None]
    k = np.arange(K)[None, :]
    A_abs = S[hop_elems * r + start_elems + k].astype(np.float32)
    P_abs = A_abs @ Bm.astype(np.float32).T
    # rel hypothesis: XOR term = row index in group (+ base offset), applied to the chunk bits of the un-swizzled address
    A_rel = np.zeros((128, K), dtype=np.float32)
    for rr in range(128):
        for kk in range(K):
            logical = 2 * (start_elems + (rr % 8) * hop_elems + (rr // 8) * 8 * hop_elems + kk)
            x = ((rr % 8) + (((2 * start_elems) >> 7) & 7 if base_offset_rule else 0)) & mask
            phys = logical ^ (x << 4)
            A_rel[rr, kk] = float(h[phys // 2])
    P_rel = A_rel @ Bm.astype(np.float32).T
    return float(np.abs(D - P_abs).max()), float(np.abs(D - P_rel).max())


def chunk_major_test(hop_chunks, n_rows_store, start_row, K, N, rng):
    """no-swizzle K-major with SBO = 128: chunk c (8 elements) of block-row r lives at 16 r + LBO c, LBO = 16 R.
    A(r, k) = S[hop*(r + start_row) + k] for k < hop = 8*hop_chunks; k >= hop continues in the next block-row."""
    R = n_rows_store
    hop = 8 * hop_chunks
    S = rng.integers(-3, 4, size=R * hop).astype(np.float16)
    img = np.zeros(IMG, dtype=np.uint8)
    h = img.view(np.float16)
    for p in range(S.size):
        r, c, e = p // hop, (p % hop) // 8, p % 8
        h[(16 * r + 16 * R * c) // 2 + e] = S[p]
    Bm = rng.integers(-3, 4, size=(N, K)).astype(np.float16)
    b_starts = put_b(img, Bm)
    mmas = []
    for ks in range(K // 16):
        k0 = 16 * ks
        row_shift, c0 = divmod(k0 // 8, hop_chunks)
        if hop_chunks == 1:
            # both chunks of the k-step are consecutive block rows: LBO = 16 (next row), i.e. a plain linear signal
            st, lbo = 16 * (start_row + k0 // 8), 16
        else:
            st, lbo = 16 * (start_row + row_shift) + 16 * R * c0, 16 * R
        mmas.append((s64(desc(st, lbo, 128, 0)), s64(desc(b_starts[ks], 16, 1024, 2)), 0, int(ks > 0), idesc_f16(128, N)))
    D, _ = run(img, mmas, 32 * ((N + 31) // 32))
    D = D[:, :N]
    r = np.arange(128)[:, None]
    k = np.arange(K)[None, :]
    A = S[hop * (r + start_row) + k].astype(np.float32)
    return float(np.abs(D - A @ Bm.astype(np.float32).T).max())


def timing(mode, N, n_mma, reps, a_stride_bytes):
    layout, pitch, _ = SW[mode]
    img = np.zeros(IMG, dtype=np.uint8)
    img.view(np.float16)[:] = np.float16(1.0)
    mmas = []
    for i in range(n_mma):
        st = (i * a_stride_bytes) % (32 * 1024)
        if mode == "none":
            a = desc(st, 16 * 1024 // 8, 128, 0)
        else:
            a = desc(st, 16, 8 * pitch, layout)
        mmas.append((s64(a), s64(desc(B_OFF + (i % 4) * 32, 16, 1024, 2)), (i % 4) * 128 if False else 0, 1, idesc_f16(128, N)))
    _, cyc = run(img, mmas, 32, reps=reps)
    return cyc / (reps * n_mma)

def window_timing(pattern, n_mma=60, reps=50, N=48):
    """cycles per MMA for accumulator-window patterns (all operands no-swizzle chunk-major, as the VQT level kernels):
    same: every MMA into columns [0, N); shift8: window start 8 * (i // 3) (the decimator's overlapping windows);
    shift16 / shift48: start 16 / 48 * (i // 3); two: alternate two disjoint windows; fb: alternate N = 48 / 32 windows."""
    img = np.zeros(IMG, dtype=np.uint8)
    img.view(np.float16)[:] = np.float16(1.0)
    mmas = []
    for i in range(n_mma):
        a = desc(16 * (i % 7), 2096, 128, 0)
        b = desc(B_OFF + 16 * (i % 5), 1792, 128, 0)
        n = N
        if pattern == "same":
            col = 0
        elif pattern == "shift8":
            col = 8 * (i // 3)
        elif pattern == "shift16":
            col = 16 * (i // 3)
        elif pattern == "shift48":
            col = 48 * ((i // 3) % 8)
        elif pattern == "two":
            col = 64 * (i % 2)
        elif pattern == "fb":
            col, n = (0, 48) if i % 2 == 0 else (48, 32)
        mmas.append((s64(a), s64(b), col, 1, idesc_f16(128, n)))
    _, cyc = run(img, mmas, 32, reps=reps)
    return cyc / (reps * n_mma)


def zero_block_test(rng):
    """accumulate = 0 with an all-zero 128-byte block as BOTH operands (LBO = SBO = 0: every row and chunk aliases it)
    must clear an accumulator window of any N without touching its neighbours."""
    img = np.zeros(IMG, dtype=np.uint8)
    h = img.view(np.float16)
    S = rng.integers(-3, 4, size=4096).astype(np.float16)
    h[: S.size] = S                                   # chunk-major q = 1 signal: row r at 16 r
    Bm = rng.integers(-3, 4, size=(256, 16)).astype(np.float16)
    b_starts = put_b(img, np.concatenate([Bm] * 4, axis=1))   # K = 64 canonical tiles
    zoff = A_BYTES - 1024                              # zero block (image is zero there)
    fill = (s64(desc(0, 16, 128, 0)), s64(desc(b_starts[0], 16, 1024, 2)), 0, 0, idesc_f16(128, 256))
    res = {}
    for n in (16, 48, 128, 256):
        zero = (s64(desc(zoff, 0, 0, 0)), s64(desc(zoff, 0, 0, 0)), 64, 0, idesc_f16(128, n))
        D, _ = run(img, [fill, zero], 512 if n > 192 else 256)
        r = np.arange(128)[:, None]
        A = S[8 * r + np.arange(16)[None, :]].astype(np.float32)
        want = A @ Bm.astype(np.float32).T
        want[:, 64:min(256, 64 + n)] = 0
        got = D[:, :256]
        if n + 64 > 256:
            assert np.abs(D[:, 256:64 + n]).max() == 0
        res[n] = float(np.abs(got - want).max())
    return res


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    print("== cycles per MMA by accumulator-window pattern (N = 48, chunk-major operands):",
          {p: round(window_timing(p), 1) for p in ("same", "shift8", "shift16", "shift48", "two", "fb")})
    print("== same, by MMAs per commit:", {n: round(window_timing("shift8", n_mma=n), 1) for n in (3, 10, 30, 60, 120)})
    print("== zeroing MMA with an aliased zero block (LBO = SBO = 0): max|D - expected| per N:", zero_block_test(np.random.default_rng(1)))
    if only == "all":
        rng = np.random.default_rng(0)
        print("== Toeplitz views through swizzled K-major descriptors: max|D - prediction| (abs-address rule, row-relative rule)")
        for mode, hop in (("sw128", 64), ("sw64", 32), ("sw32", 16)):
            for start in (0, 8, hop, hop + 24, 3 * hop + 8, 5 * hop):
                for rule in (0, 1):
                    try:
                        ea, er = toeplitz_test(mode, hop, start, 128, 32, rule, rng)
                        print(f"{mode} hop={hop} start={start:4d} base_offset_field={'(start>>7)&7' if rule else '0':12s}: abs {ea:8.1f}  rel {er:8.1f}")
                    except Exception as e:  # noqa: BLE001
                        print(f"{mode} start={start} rule={rule}: FAILED {e}")
        print("== no-swizzle K-major, SBO=128 (rows at 16-byte pitch), LBO = 16*R: max|D - prediction|")
        for hop_chunks, R, start_row, K in ((1, 2048, 0, 64), (1, 2048, 5, 96), (4, 256, 0, 96), (4, 256, 3, 96), (8, 200, 1, 128), (2, 515, 2, 64)):
            try:
                e = chunk_major_test(hop_chunks, R, start_row, K, 48, rng)
                print(f"hop={8*hop_chunks:3d} R={R} start_row={start_row} K={K}: {e:8.1f}")
            except Exception as ex:  # noqa: BLE001
                print(f"hop_chunks={hop_chunks}: FAILED {ex}")
        print("== cycles per MMA (128 x N x 16, fp16), one CTA, 64 MMAs per commit, 50 reps")
        for mode in ("sw128", "sw64", "sw32", "none"):
            for N in (16, 32, 48, 64, 96, 128, 256):
                for stride in (0, 32, 2048):
                    try:
                        c = timing(mode, N, 64, 50, stride)
                        print(f"{mode:6s} N={N:3d} a_stride={stride:5d}: {c:7.1f} clk/MMA")
                    except Exception as ex:  # noqa: BLE001
                        print(f"{mode} N={N}: FAILED {ex}")
