"""CPU model of the index arithmetic shared by the double-march adjoint kernels (no GPU needed).

The chunking of dimension 3 is decided in three places that must agree: the host plan
(`sg_adjoint_march2_plan`, csrc/sg_fast_adjoint.cu: number of chunks and worst-case chunk length G3 = row stride of
the partials), the march kernels (`sg_m2_chunk_len`, csrc/sg_fast_adjoint.cuh: chunks adapt to the spans that hold
samples) and the post kernel's chunk look-up (csrc/sg_adjoint_post2.cuh).  This file restates the three pieces in
Python and checks exhaustively, for small sizes, that every control plane of the support is found in exactly the
chunks that wrote it, in a row inside the allocated stride, and in at most two chunks.
"""
import itertools


def host_plan(nsp3, P, want_chunks):
    chunks = max(1, want_chunks)
    chunks = min(chunks, max(1, nsp3 // max(2, P)))
    G3 = max(max(2, P), -(-nsp3 // chunks))
    return chunks, G3


def chunk_len(nact, P, chunks3):                      # sg_m2_chunk_len
    return max(max(P, 1), -(-nact // chunks3))


def written_rows(sf, sl, P, chunks3):
    """chunk -> {control index i3 (1-based): row l3} as the march kernels write them."""
    G3e = chunk_len(sl - sf + 1, P, chunks3)
    out = {}
    for c in range(chunks3):
        s_lo = sf + c * G3e
        s_hi = min(s_lo + G3e, sl + 1)
        if s_lo >= s_hi:
            continue                                  # the CTA returns at once
        out[c] = {i3: i3 - (s_lo - P) for i3 in range(s_lo - P, s_hi)}
    return out


def post_lookup(i3, sf, sl, P, chunks3):
    """(chunk, row) pairs the post kernel reads for control index i3."""
    G3e = chunk_len(sl - sf + 1, P, chunks3)
    nch = (sl - sf + 1 + G3e - 1) // G3e
    c_lo = (i3 - sf) // G3e if i3 >= sf else 0
    c_hi = min((i3 - sf + P) // G3e, nch - 1)
    return [(c, i3 - (sf + c * G3e - P)) for c in range(c_lo, c_hi + 1)]


def test_post_kernel_reads_exactly_what_the_march_kernels_wrote():
    checked = 0
    for P, nsp3, want in itertools.product((1, 2, 3), range(1, 34), (1, 2, 3, 5, 8, 13, 40)):
        chunks3, G3 = host_plan(nsp3, P, want)
        assert G3 >= P and chunks3 >= 1
        c3 = nsp3 + P                                 # control indices of dimension 3
        for sf in range(P + 1, c3 + 1):               # first / last span that holds samples (1-based spans P+1..c3)
            for sl in range(sf, c3 + 1):
                G3e = chunk_len(sl - sf + 1, P, chunks3)
                assert P <= G3e <= G3, (P, nsp3, want, sf, sl)
                wrote = written_rows(sf, sl, P, chunks3)
                assert wrote and max(wrote) < chunks3
                for rows in wrote.values():
                    assert all(0 <= l < G3 + P for l in rows.values())          # inside the allocated row stride
                for i3 in range(sf - P, sl + 1):      # the support of the (slab of the) grid
                    expect = sorted((c, rows[i3]) for c, rows in wrote.items() if i3 in rows)
                    got = sorted(post_lookup(i3, sf, sl, P, chunks3))
                    assert got == expect, (P, nsp3, want, sf, sl, i3, got, expect)
                    assert 1 <= len(got) <= 2
                    checked += 1
    assert checked > 100000


def test_tiles_of_dimension_2_cover_every_control_row_once_or_twice():
    """Post kernel: tile t owns control rows t*G2 .. t*G2+G2-1 (t = 0..tiles2, the last one only the halo rows); a row
    gets slot q of its own tile and slot q+G2 of the previous tile (q < P).  Every (tile, slot) the march kernel writes
    for a real control row must be read exactly once."""
    G2 = 4
    for P in (1, 2, 3):
        S = G2 + P
        for nsp2 in range(1, 40):
            c2 = nsp2 + P
            tiles2 = -(-nsp2 // G2)
            written = {(t, s) for t in range(tiles2) for s in range(S) if t * G2 + s < c2}   # slot s of tile t = row t*G2+s
            read = []
            for t in range(tiles2 + 1):
                for q in range(G2):
                    i2 = t * G2 + q
                    if i2 >= c2:
                        continue
                    if t < tiles2:
                        read.append((t, q))
                    if t >= 1 and q < P:
                        read.append((t - 1, q + G2))
            assert len(read) == len(set(read))
            assert set(read) == written, (P, nsp2)
            assert (tiles2 + 1) * G2 >= c2            # every control row belongs to some CTA
