"""The N>1 path on CPU: two gloo ranks, doc-range shards with GLOBAL statistics, all-gather of
the per-shard lists and the merge.  The CUDA scoring / merge kernels cannot run here, so the
oracle stands in for them as injected callables; what is under test is the host logic of
probing_rag_b200/sharding.py (ranges, integer all-reduce of df / token counts, rank-major
gather layout) -- the merged result must equal the single-index oracle result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bm25_oracle as bo
from probing_rag_b200 import sharding, synth

N_DOCS, VOCAB, NQ, K = 6000, 1 << 12, 64, 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_merge(gs, gd):
    s, d = bo.merge_topk(gs.numpy(), gd.numpy(), gs.shape[2])
    return torch.from_numpy(s), torch.from_numpy(d)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(N_DOCS, rank, world)
        toks, lens = synth.corpus_np(N_DOCS, VOCAB, doc_lo=lo, doc_hi=hi)
        df_local = torch.from_numpy(np.bincount(
            np.unique(toks.astype(np.int64) * N_DOCS + np.repeat(np.arange(len(lens)), lens)) // N_DOCS,
            minlength=VOCAB))
        df, avgdl = sharding.global_stats(df_local, int(lens.sum()), N_DOCS)
        shard = bo.build_index(toks, lens, VOCAB, n_docs_global=N_DOCS, avgdl_global=avgdl,
                               df_global=df.numpy(), doc_id_base=lo)
        qi, qt = synth.queries_np(NQ, VOCAB, df.numpy())

        def local_topk(q_indptr, q_terms, k):
            s, d = bo.retrieve_batch(shard, q_indptr.numpy(), q_terms.numpy(), k)
            return torch.from_numpy(s), torch.from_numpy(d)

        sb = sharding.ShardedBM25(local_topk=local_topk, merge=_oracle_merge)
        s, d = sb.topk(torch.from_numpy(qi), torch.from_numpy(qt), K)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), s=s.numpy(), d=d.numpy(), df=df.numpy(), avgdl=avgdl)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_doc_shards_equal_single_index(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    toks, lens = synth.corpus_np(N_DOCS, VOCAB)
    full = bo.build_index(toks, lens, VOCAB)
    qi, qt = synth.queries_np(NQ, VOCAB, full["df"])
    ws, wd = bo.retrieve_batch(full, qi, qt, K)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(got["df"], full["df"])
        assert float(got["avgdl"]) == full["avgdl"]
        assert np.array_equal(got["d"], wd), f"rank {r}: merged doc ids differ from the single index"
        assert np.array_equal(got["s"], ws), f"rank {r}: merged scores differ from the single index"


def test_shard_ranges_cover_corpus():
    for n, w in ((10, 3), (21_015_324, 8), (5, 8), (0, 2)):
        r = [sharding.shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)
