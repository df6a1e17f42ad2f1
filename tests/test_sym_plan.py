"""The work decomposition of the symmetric dense preconditioner apply (dpgo_sym_plan, host only -- the very plan
phase_dense_sym / k_pack_sym / phase_pz run on), checked on the CPU:

  * structure: segment-major chunk order, contiguous non-empty CTA runs, consecutive CTAs per segment, packed offsets;
  * arithmetic: a NumPy emulation of the kernels with the SAME tables -- pack the upper triangle chunk-major, per CTA
    run accumulate the direct product into its panel slot and write the transposed product of every chunk into its
    segment slot, then sum slots as phase_pz does -- reproduces V @ P for a random symmetric P (the reference operator
    is (Q + 0.1 I)^-1, src/QuadraticProblem.cpp:75-87), including ragged sizes (N not a multiple of 8 or of 480)."""
import ctypes as C

import numpy as np
import pytest

SEG = 480


def get_plan(N, grid, cost=0.0):
    from dpo_b200 import _capi
    lib = _capi.load_library()
    nseg, nch = C.c_int(), C.c_int()
    assert lib.dpgo_sym_plan_sizes(N, C.byref(nseg), C.byref(nch)) == 0
    segptr = np.zeros(nseg.value + 1, np.int32)
    cut = np.zeros(grid + 1, np.int32)
    cfirst = np.zeros(nseg.value, np.int32)
    ccount = np.zeros(nseg.value, np.int32)
    off = np.zeros(nch.value + 1, np.int64)
    rc = lib.dpgo_sym_plan(N, grid, C.c_double(cost), _capi.iptr(segptr), _capi.iptr(cut), _capi.iptr(cfirst),
                           _capi.iptr(ccount), off.ctypes.data_as(C.POINTER(C.c_int64)))
    return rc, nseg.value, nch.value, segptr, cut, cfirst, ccount, off


def chunks_of(N):
    """(J, g, col_lo, s1) of every chunk in segment-major order."""
    out = []
    for J in range((N + SEG - 1) // SEG):
        s0, s1 = J * SEG, min(N, (J + 1) * SEG)
        for g in range((s1 + 7) // 8):
            out.append((J, g, max(s0, 8 * g), s1))
    return out


@pytest.mark.parametrize("N,grid", [(2048, 148), (2500, 148), (6644, 148), (10000, 148), (10000, 132), (20000, 148),
                                    (2050, 8), (3002, 1)])
def test_plan_structure(N, grid):
    rc, nseg, nch, segptr, cut, cfirst, ccount, off = get_plan(N, grid)
    assert rc == 0
    ch = chunks_of(N)
    assert nch == len(ch) and nseg == (N + SEG - 1) // SEG
    assert segptr[0] == 0 and segptr[-1] == nch
    for J in range(nseg):
        assert segptr[J + 1] - segptr[J] == (min(N, (J + 1) * SEG) + 7) // 8
    # runs: contiguous, cover everything, none empty, nearly equal chunk counts (the cost model is count-dominated)
    assert cut[0] == 0 and cut[-1] == nch and np.all(np.diff(cut) >= 1)
    counts = np.diff(cut)
    assert counts.max() <= 1.35 * counts.mean() + 2
    # per segment the touching CTAs are consecutive and cfirst/ccount describe exactly them
    for J in range(nseg):
        touch = [b for b in range(grid) if cut[b] < segptr[J + 1] and cut[b + 1] > segptr[J]]
        assert touch == list(range(cfirst[J], cfirst[J] + ccount[J]))
    # packed layout: 8 rows x (width rounded up to 8, + 4) doubles per chunk, 16-byte aligned, no overlap
    for lin, (J, g, col_lo, s1) in enumerate(ch):
        pitch = ((s1 - col_lo + 7) & ~7) + 4
        assert off[lin + 1] - off[lin] == 8 * pitch
        assert pitch % 16 in (4, 12)                      # row pitch = 32 or 96 bytes mod 128: conflict-free tile reads
    assert off[0] == 0
    # every 8x8 tile of the upper triangle (tile row <= tile column) is covered by exactly one chunk
    ntile = (N + 7) // 8
    cover = np.zeros((ntile, ntile), np.int32)
    for (J, g, col_lo, s1) in ch:
        cover[g, col_lo // 8:(s1 + 7) // 8] += 1
    assert np.array_equal(cover, np.triu(np.ones((ntile, ntile), np.int32)))


@pytest.mark.parametrize("N", [1000, 2047, 2049])
def test_no_plan_for_small_or_odd(N):
    assert get_plan(N, 148)[0] != 0


def test_no_plan_when_more_ctas_than_chunks():
    assert get_plan(2048, 4096)[0] != 0


@pytest.mark.parametrize("N,grid,r", [(2048, 148, 5), (2500, 148, 5), (2090, 37, 3), (3002, 148, 4)])
def test_emulated_apply_matches_dense_product(N, grid, r):
    rc, nseg, nch, segptr, cut, cfirst, ccount, off = get_plan(N, grid)
    assert rc == 0
    rng = np.random.default_rng(N + grid)
    A = rng.standard_normal((N, N))
    P = A + A.T
    V = rng.standard_normal((r, N))
    ch = chunks_of(N)
    # --- k_pack_sym: chunk-major packed upper triangle (zero fill)
    ppack = np.zeros(off[-1])
    for lin, (J, g, col_lo, s1) in enumerate(ch):
        width = s1 - col_lo
        pitch = ((width + 7) & ~7) + 4
        tile = np.zeros((8, pitch))
        rows = min(8, N - 8 * g)
        tile[:rows, :width] = P[8 * g:8 * g + rows, col_lo:s1]
        ppack[off[lin]:off[lin + 1]] = tile.ravel()
    # --- phase_dense_sym: per CTA run, direct product into its panel slot, transposed product into segment slot J
    nslot = int(ccount.max())
    part = np.zeros((nslot, r, N))
    t2 = np.zeros((nseg, r, N))
    written = np.zeros((nslot, nseg), bool)
    for b in range(grid):
        lin = cut[b]
        while lin < cut[b + 1]:
            J = int(np.searchsorted(segptr, lin, side="right") - 1)
            run_end = min(cut[b + 1], segptr[J + 1])
            s0, s1 = J * SEG, min(N, (J + 1) * SEG)
            D1 = np.zeros((r, SEG))
            for l in range(lin, run_end):
                _, g, col_lo, _ = ch[l]
                width = s1 - col_lo
                pitch = ((width + 7) & ~7) + 4
                tile = ppack[off[l]:off[l + 1]].reshape(8, pitch)[:, :width]
                rows = min(8, N - 8 * g)
                g0 = 8 * g
                Vg = np.zeros((r, 8))
                Vg[:, :rows] = V[:, g0:g0 + rows]
                # the 8x8 tiles of the chunk: direct for ctile >= g0, transposed for ctile > g0 (diagonal tile once)
                D1[:, col_lo - s0:s1 - s0] += Vg @ tile
                cols = np.arange(col_lo, s1)
                right = cols >= g0 + 8 if col_lo == g0 else np.ones(width, bool)
                T = V[:, cols[right]] @ tile[:, right].T
                t2[J, :, g0:g0 + rows] = T[:, :rows]
            slot = b - cfirst[J]
            assert 0 <= slot < ccount[J] and not written[slot, J]
            part[slot, :, s0:s1] = D1[:, :s1 - s0]
            written[slot, J] = True
            lin = run_end
    # --- phase_pz: ccount[J] panel slots of the column's segment + the transposed slots jt0 .. nseg-1
    Z = np.zeros((r, N))
    for col in range(N):
        J = col // SEG
        jt0 = (col & ~7) // SEG
        Z[:, col] = part[:ccount[J], :, col].sum(axis=0) + t2[jt0:, :, col].sum(axis=0)
    ref = V @ P
    assert np.linalg.norm(Z - ref) <= 1e-12 * np.linalg.norm(ref)
    for J in range(nseg):
        assert written[:ccount[J], J].all()
