"""The N > 1 path on real hardware: one process per GPU over NCCL (torchrun layout), doc-range shards with the
threshold exchange between launches, all-gather + merge -- the merged lists must be bit-identical to the
single-index lists AND to the CPU oracle.  Skipped on a box with one GPU (the gloo tests cover the host logic
there, tests/test_gpu_bm25.py the exchange protocol itself)."""
import json
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_DOCS, VOCAB, NQ = 300_000, 1 << 18, 512


TEXT_QUERIES = ["paris tower capital", "river bridge london", "", "probing hidden states of language models"]


def _texts():
    rng = np.random.default_rng(17)
    words = ("retrieval augmented generation probing language model hidden state wikipedia passage question answer paris "
             "france capital city river tower london england bridge").split()
    return [" ".join(rng.choice(words, size=int(rng.integers(3, 30)))) for _ in range(5000)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import bench
        from probing_rag_b200.sharding import ShardedBM25
        dev = torch.device("cuda", rank)
        gi, qi, qt = bench.build_workload(N_DOCS, VOCAB, NQ, dev, rank, world)
        # many launches per call so the exchange really runs; every rank the same number
        gi.set_tuning(subs_per_item=2, docs_per_launch=16384, min_items=1, items_per_warp=1)
        d_qi, d_qt = torch.from_numpy(qi).to(dev), torch.from_numpy(qt).to(dev)
        res = {}
        for name, mode, list_rounds in (("p2p", "p2p", 4), ("p2p_live_only", "p2p", 0), ("exchange", "allreduce", 0),
                                        ("plain", None, 0), ("p2p_again", "p2p", 100)):
            if name == "p2p_again":   # ... this time with the kernel variant a 65,536-query batch runs, and a union
                gi.set_tuning(batch_variant=2)      # bound behind every launch
            sb = ShardedBM25(gi, exchange=mode, max_queries=NQ, list_rounds=list_rounds)
            assert sb.exchange == mode and sb.list_rounds == (list_rounds if mode == "p2p" else 0)
            for k in (10, 100, 10):                 # an odd number of calls: the next p2p instance starts on the other parity
                s, d = sb.topk(d_qi, d_qt, k)
                hs, hd, h2d, d2h = sb.topk_host(qi, qt, k)
                torch.cuda.synchronize()
                assert np.array_equal(hs, s.cpu().numpy()) and np.array_equal(hd, d.cpu().numpy())
                assert d2h == NQ * k * 8
                res[f"{name}_s{k}"], res[f"{name}_d{k}"] = hs, hd
            sb.close()
        # the persisted text-level retriever, loaded sharded: every rank cuts its doc range out of the saved index
        from probing_rag_b200 import BM25Retriever
        if rank == 0:
            BM25Retriever.from_texts(iter(_texts()), similarity_top_k=4, device=dev, persist_dir=os.path.join(out_dir, "text_ix"))
        dist.barrier()
        tr = BM25Retriever.from_persist_dir(os.path.join(out_dir, "text_ix"), device=dev, sharded=True, max_queries=64)
        assert tr.index.exchange == "p2p" and tr.index.index.n_docs < len(_texts())
        rows = [[(n.node.id_, n.score, n.text) for n in one] for one in tr.retrieve_batch(TEXT_QUERIES)]
        with open(os.path.join(out_dir, f"text_r{rank}.json"), "w") as f:
            json.dump(rows, f)
        tr.index.close()
        res["launches"] = np.array([gi.num_launches(NQ, 10)])
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), **res)
        with pytest.raises(ValueError):                 # bad term ids raise on the sharded path too
            ShardedBM25(gi, exchange="allreduce").topk(d_qi, torch.full_like(d_qt, VOCAB), 10)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4])
def test_nccl_doc_shards_equal_single_index_and_oracle(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    import torch.multiprocessing as mp

    import bench
    from oracle import c_oracle as co
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    gi, qi, qt = bench.build_workload(N_DOCS, VOCAB, NQ, torch.device("cuda", 0))
    host = {"data": gi.weights.cpu().numpy(), "indices": gi.doc_ids.cpu().numpy(), "indptr": gi.indptr.cpu().numpy(),
            "num_docs": gi.n_docs}
    d_qi, d_qt = torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda()
    from probing_rag_b200 import BM25Retriever
    single = BM25Retriever.from_persist_dir(str(tmp_path / "text_ix"))
    want = [[[n.node.id_, n.score, n.text] for n in one] for one in single.retrieve_batch(TEXT_QUERIES)]
    assert all(len(one) == 4 and all(x[2] for x in one) for one in want)
    for r in range(world):
        assert json.load(open(tmp_path / f"text_r{r}.json")) == want, f"rank {r}: sharded text retrieval differs"
    for k in (10, 100):
        s1, d1 = gi.topk(d_qi, d_qt, k)
        os_, od = co.retrieve_batch(host, qi, qt, k, n_threads=min(16, os.cpu_count() or 1))
        assert np.array_equal(d1.cpu().numpy(), od) and np.array_equal(s1.cpu().numpy(), os_)
        for r in range(world):
            got = np.load(tmp_path / f"r{r}.npz")
            assert int(got["launches"][0]) > 2
            for name in ("p2p", "p2p_live_only", "exchange", "plain", "p2p_again"):
                assert np.array_equal(got[f"{name}_d{k}"], od), f"rank {r} {name}: merged doc ids differ (k={k})"
                assert np.array_equal(got[f"{name}_s{k}"], os_), f"rank {r} {name}: merged scores differ (k={k})"
