"""CPU model of the fused IVF-PQ search's candidate funnel (semtools_b200/csrc/ivfpq.cu,
ivf_adc_finish_kernel): 32-code chunks dealt round-robin over CTAs first, then warps; per-warp
top-64 -> per-CTA top-256 -> global top-`rerank`.  Checks the index arithmetic (every code is
scanned exactly once) and that the funnel is lossless for the best `rerank` codes even when
they are CONTIGUOUS in code order (a tight cluster inside one inverted list), which is what the
CTA-first dealing is for."""
import numpy as np
import pytest

N_CTAS, WARPS, KEEP_WARP, KEEP_CTA = 32, 16, 64, 256


def deal(total):
    """(cta, warp) -> code positions v, exactly as the kernel's loop."""
    owner = {}
    n_chunks = (total + 31) // 32
    for b in range(N_CTAS):
        for w in range(WARPS):
            g = b + N_CTAS * w
            vs = []
            while g * 32 < total:
                vs.extend(range(g * 32, min(g * 32 + 32, total)))
                g += N_CTAS * WARPS
            owner[(b, w)] = np.asarray(vs, dtype=np.int64)
    return owner, n_chunks


def funnel(scores, rerank):
    owner, _ = deal(len(scores))
    survivors = []
    for b in range(N_CTAS):
        cta = []
        for w in range(WARPS):
            v = owner[(b, w)]
            if len(v):
                o = v[np.lexsort((v, -scores[v]))][:KEEP_WARP]        # score desc, position asc
                cta.append(o)
        if cta:
            v = np.concatenate(cta)
            survivors.append(v[np.lexsort((v, -scores[v]))][:KEEP_CTA])
    v = np.concatenate(survivors)
    return v[np.lexsort((v, -scores[v]))][:rerank]


@pytest.mark.parametrize("total", [0, 1, 31, 32, 33, 16_384, 16_385, 62_395, 200_003])
def test_every_code_is_scanned_exactly_once(total):
    owner, _ = deal(total)
    allv = np.concatenate([v for v in owner.values()]) if total else np.zeros(0, np.int64)
    assert len(allv) == total and np.array_equal(np.sort(allv), np.arange(total))


@pytest.mark.parametrize("case", ["uniform", "contiguous_cluster", "two_lists", "ties"])
def test_funnel_is_lossless_for_the_best_rerank(case):
    rng = np.random.default_rng(3)
    total, rerank = 62_395, 512
    s = rng.standard_normal(total).astype(np.float32)
    if case == "contiguous_cluster":
        s[20_000:20_700] += 10.0                                       # the best 700 codes are contiguous
    elif case == "two_lists":
        s[1_000:1_300] += 10.0; s[40_000:40_400] += 10.0
    elif case == "ties":
        s[5_000:9_000] = 7.0                                           # 4000 equal scores: position order decides
    want = np.arange(total)[np.lexsort((np.arange(total), -s))][:rerank]
    got = funnel(s, rerank)
    assert np.array_equal(got, want)
