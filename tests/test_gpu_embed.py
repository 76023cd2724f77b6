"""GPU parity: K3 (gather + mean-pool + L2-normalise) through the C ABI is
BIT-IDENTICAL to the oracle's restatement of model2vec-rs pool_ids."""
import os

import numpy as np
import pytest

import oracle
from semtools_b200 import capi

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_golden_pool_cases(ctx):
    z = np.load(os.path.join(G, "pool_small.npz"))
    for variant in ["plain", "weights", "mapping", "both", "nonorm"]:
        w = z["weights"] if variant in ("weights", "both") else None
        m = z["mapping"] if variant in ("mapping", "both") else None
        t = capi.Table(ctx, z["E"], w, m, normalize=(variant != "nonorm"))
        out = capi.embed(ctx, t, z["offsets"], z["ids"])
        assert np.array_equal(bits(out), bits(z[f"out_{variant}"])), variant
        t.close()


def synth_batch(rng, n_lines, V, mu=2.5, sigma=0.8, zipf=1.1):
    """SURVEY 8d token-id generator: lengths ~ clamp(round(LogNormal)), ids ~ Zipf."""
    T = np.clip(np.round(rng.lognormal(mu, sigma, n_lines)), 0, 2048).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(T)]).astype(np.uint64)
    ids = (rng.zipf(zipf, int(T.sum())) - 1) % V
    return offsets, ids.astype(np.uint32)


def test_random_batch_bit_exact_and_appends_to_corpus(ctx):
    rng = np.random.default_rng(0x5E117001)
    V = 50_000
    E = (rng.standard_normal((V, 256)) * 0.1).astype(np.float32)
    offsets, ids = synth_batch(rng, 20_000, V)
    assert (np.diff(offsets.astype(np.int64)) == 0).any()          # includes empty lines
    t = capi.Table(ctx, E)
    exp = oracle.embed_csr(E, offsets, ids)
    out = capi.embed(ctx, t, offsets, ids)
    assert np.array_equal(bits(out), bits(exp))
    c = capi.Corpus(ctx, 16)
    c.append(np.ones((3, 256), dtype=np.float32))
    capi.embed(ctx, t, offsets, ids, out=False, append_to=c)
    assert len(c) == 3 + 20_000
    assert np.array_equal(bits(c.read(3)), bits(exp))
    # downstream: searching the GPU-embedded corpus == oracle over oracle embeddings
    q = exp[17]
    r, d = oracle.search_rows(np.concatenate([np.ones((3, 256), np.float32), exp]), q, top_k=10)
    hits = c.search(q, top_k=10)
    assert hits["row"].tolist() == [int(x) for x in r]
    assert np.array_equal(hits["distance"], d)


def test_long_lines_and_query_truncation_lengths(ctx):
    rng = np.random.default_rng(3)
    V = 3000
    E = (rng.standard_normal((V, 256)) * 0.1).astype(np.float32)
    w = rng.uniform(0.1, 2.0, V).astype(np.float32)
    lens = [2048, 512, 513, 1, 0, 9, 16, 17]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(0, V, int(offsets[-1])).astype(np.uint32)
    t = capi.Table(ctx, E, weights=w)
    assert np.array_equal(bits(capi.embed(ctx, t, offsets, ids)), bits(oracle.embed_csr(E, offsets, ids, w)))


def test_out_of_range_token_fails_and_appends_nothing(ctx):
    E = np.ones((10, 256), dtype=np.float32)
    t = capi.Table(ctx, E)
    c = capi.Corpus(ctx, 8)
    with pytest.raises(capi.StbError) as e:
        capi.embed(ctx, t, [0, 2, 3], [1, 10, 2], out=False, append_to=c)
    assert e.value.status == capi.STB_ERR_RANGE
    assert len(c) == 0
    capi.embed(ctx, t, [0, 2, 3], [1, 9, 2], out=False, append_to=c)
    assert len(c) == 2


def test_device_resident_embed_entry_point(ctx):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(4)
    V = 4000
    E = (rng.standard_normal((V, 256)) * 0.1).astype(np.float32)
    offsets, ids = synth_batch(rng, 3000, V)
    t = capi.Table(ctx, E)
    dev = torch.device("cuda:0")
    off_d = torch.from_numpy(offsets.view(np.int64)).to(dev)
    ids_d = torch.from_numpy(ids.view(np.int32)).to(dev)
    out_d = torch.zeros((3000, 256), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    capi.embed_dev(ctx, t, off_d.data_ptr(), ids_d.data_ptr(), 3000, out_d.data_ptr())
    capi.embed_status(ctx)
    assert np.array_equal(bits(out_d.cpu().numpy()), bits(oracle.embed_csr(E, offsets, ids)))
    ids_d[5] = V + 3                                   # out of range -> sticky flag, reported once
    torch.cuda.synchronize()
    capi.embed_dev(ctx, t, off_d.data_ptr(), ids_d.data_ptr(), 3000, out_d.data_ptr())
    with pytest.raises(capi.StbError) as e:
        capi.embed_status(ctx)
    assert e.value.status == capi.STB_ERR_RANGE
    capi.embed_status(ctx)                             # cleared
