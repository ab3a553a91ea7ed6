"""Doc-range sharding of the BM25 index over the GPUs of one box.

The reference is single-process (`exp_rag.py:242`: one BM25Retriever over the whole corpus);
BM25 shards naturally because a document's score depends only on its own postings and on the
GLOBAL constants N, avgdl, df.  One process per GPU (torchrun):

    rank g owns docs [g*ceil(N/G), (g+1)*ceil(N/G))                       shard_range
    df and the token count are summed over ranks once, at build time       global_stats
    every rank scores the whole query batch against its shard              BM25Index.topk
      between the launches of that call the per-query lower bounds of the
      final k-th score are raised to the best any rank knows               exchange_thresholds
    [B,k] (score, doc id) lists are all-gathered                           gather_lists
    and merged in the canonical order (score desc, doc id asc)             pr_topk_merge

The merged lists equal the single-index lists bit for bit: each document is scored on exactly
one rank with the same weights and the same fp32 summation order, and a bound found on one
rank (k documents there score at least that much) can never exclude a document of the global
top-k on another (the global k-th score is >= every rank's local k-th score).  The collective plumbing
below is backend-agnostic (NCCL on the GPUs, gloo in the CPU tests); the scoring and the merge
are CUDA only (`local_topk` / `merge` are injectable so the host logic can be tested with the
oracle standing in as the checker).
"""
from __future__ import annotations

import ctypes
from typing import Callable

import torch
import torch.distributed as dist

from . import _lib


def shard_range(n_docs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous doc-id range [lo, hi) of `rank`; the last ranks may be short or empty."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    per = -(-n_docs // world)
    return min(rank * per, n_docs), min((rank + 1) * per, n_docs)


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def global_stats(df_local: torch.Tensor, n_tokens_local: int, n_docs_global: int, group=None):
    """(df_global i64[V], avgdl float): the statistics bm25s computes over the whole corpus
    (App. A.4), from per-shard counts.  Integer sums, so the result does not depend on the
    number of shards."""
    df = df_local.to(torch.int64).clone()
    n_tok = torch.tensor([int(n_tokens_local)], dtype=torch.int64, device=df.device)
    if _world(group) > 1:
        dist.all_reduce(df, group=group)
        dist.all_reduce(n_tok, group=group)
    return df, float(n_tok.item()) / float(max(n_docs_global, 1))


def gather_lists(scores: torch.Tensor, ids: torch.Tensor, group=None, out=None):
    """[B,k] per-rank lists -> ([G,B,k] scores, [G,B,k] ids), rank-major."""
    world = _world(group)
    if out is None:
        out = (torch.empty((world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device),
               torch.empty((world,) + tuple(ids.shape), dtype=ids.dtype, device=ids.device))
    if world == 1:
        out[0][0].copy_(scores)
        out[1][0].copy_(ids)
    else:
        # concatenation along dim 0 ([G*B, k] view): the form every backend accepts
        flat = (world * scores.shape[0],) + tuple(scores.shape[1:])
        dist.all_gather_into_tensor(out[0].view(flat), scores.contiguous(), group=group)
        dist.all_gather_into_tensor(out[1].view(flat), ids.contiguous(), group=group)
    return out


def exchange_thresholds(theta: torch.Tensor, group=None) -> None:
    """theta f32[B] (-1 = no bound yet) <- elementwise maximum over the ranks, in place.  262 KB at 64k
    queries: a latency-sized all-reduce, enqueued behind the launch that produced the bounds."""
    if _world(group) > 1:
        dist.all_reduce(theta, op=dist.ReduceOp.MAX, group=group)


class PeerThresholds:
    """Per-query score bounds that the GPUs of one box raise in EACH OTHER's memory while they score
    (include/probing_rag.h, pr_index_set_peer_thetas): this rank's array of 2 x capacity floats in
    IPC-shareable device memory, every other rank's array opened over NVLink, and the pointer table the
    kernels index.  One process per GPU; the 64-byte IPC handles travel through `all_gather_object`."""

    def __init__(self, index, capacity: int, group=None):
        L = _lib.lib()
        self.index, self.group = index, group
        self.capacity = int(capacity)
        self.device = index.device
        dev_no = self.device.index or 0
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world - 1 > _lib.PR_MAX_PEERS:
            raise ValueError(f"at most {_lib.PR_MAX_PEERS + 1} ranks can share thresholds")
        self._local = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        self._peers: list[ctypes.c_void_p] = []
        with torch.cuda.device(self.device):
            _lib.check(L.pr_peer_alloc(dev_no, 2 * self.capacity * 4, ctypes.byref(self._local), handle))
            handles = [None] * world
            dist.all_gather_object(handles, (rank, handle.raw), group=group)
            for r, raw in handles:
                if r == rank:
                    continue
                ptr = ctypes.c_void_p()
                _lib.check(L.pr_peer_open(dev_no, raw, ctypes.byref(ptr)))
                self._peers.append(ptr)
            self._table = torch.zeros(2 * _lib.PR_MAX_PEERS, dtype=torch.int64, device=self.device)
            bases = (ctypes.c_void_p * max(len(self._peers), 1))(*[p.value for p in self._peers])
            _lib.check(L.pr_index_set_peer_thetas(index._handle, self._local, self.capacity, len(self._peers), bases,
                                                  self._table.data_ptr()))
        # nobody may raise a bound in an array that its owner has not initialised yet
        dist.barrier(group=group)

    def close(self) -> None:
        L = _lib.lib()
        if getattr(self, "_local", None) is None:
            return
        try:
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(group=self.group)          # every rank has stopped writing into the others' arrays
            with torch.cuda.device(self.device):
                L.pr_index_set_peer_thetas(self.index._handle, None, 0, 0, None, None)
                for p in self._peers:
                    L.pr_peer_close(p)
                L.pr_peer_free(self._local)
        finally:
            self._local, self._peers = None, []


class ShardedBM25:
    """This rank's shard + the exchange steps.  `local_topk(q_indptr, q_terms, k) -> (scores, ids)`
    defaults to the shard's CUDA kernel (with the threshold exchange between its launches),
    `merge([G,B,k], [G,B,k]) -> ([B,k], [B,k])` to pr_topk_merge.  Every rank returns the same
    global lists, so a `BM25Retriever(index=ShardedBM25(shard))` behaves like the single-index one."""

    doc_id_base = 0          # the merged lists carry global doc ids

    def __init__(self, index=None, group=None, local_topk: Callable | None = None,
                 merge: Callable | None = None, exchange="allreduce", max_queries: int = 65536, list_rounds: int = 4):
        """exchange: how the shards tell each other their score bounds while a call runs --
        "p2p": raised live in each other's memory by the scoring warps (NVLink peer memory, `PeerThresholds`;
               batches of up to `max_queries` queries);
        "allreduce" (or True): an all-reduce(MAX) between the launches of the call;
        None / False: not at all (every shard filters with what it found itself).
        list_rounds: with "p2p", behind each of the first `list_rounds` launches of a call (the ramp of a large
               batch, while bounds are weak and most tiles are still scanned) the shards' running score lists are
               all-gathered and every bound raised to the k-th largest score of the UNION -- stronger than the best
               single shard's k-th score that the live exchange carries.  0 = never; calls of one launch (small
               batches) never exchange."""
        if index is None and local_topk is None:
            raise ValueError("pass the shard's BM25Index or a local_topk callable")
        if exchange is True:
            exchange = "allreduce"
        if exchange not in ("p2p", "allreduce", None, False):
            raise ValueError(f"exchange must be 'p2p', 'allreduce' or None, got {exchange!r}")
        self.index = index
        self.group = group
        self._local = local_topk
        self._merge = merge
        self._gath = {}
        real = index is not None and local_topk is None and _world(group) > 1
        self.exchange = exchange if (real and exchange) else None
        self._exchange = self.exchange == "allreduce"
        self._peers = PeerThresholds(index, max_queries, group) if self.exchange == "p2p" else None
        self.list_rounds = int(list_rounds) if self.exchange == "p2p" else 0
        self._run_gath = {}
        self._max_docs = None
        if self._exchange or self.list_rounds > 0:
            # every rank must join the same number of all-reduces per call: as many as the LONGEST shard has launches
            n = torch.tensor([index.n_docs], dtype=torch.int64, device=index.device)
            dist.all_reduce(n, op=dist.ReduceOp.MAX, group=group)
            self._max_docs = int(n.item())

    def close(self) -> None:
        """Release the peer-shared threshold arrays (collective: every rank calls it)."""
        if self._peers is not None:
            self._peers.close()
            self._peers = None
            self.exchange = None

    @property
    def n_docs_global(self) -> int:
        return self.index.n_docs_global if self.index is not None else -1

    def _local_topk(self, q_indptr, q_terms, k: int):
        if self._local is not None:
            return self._local(q_indptr, q_terms, k)
        if not self._exchange:
            nq = q_indptr.numel() - 1
            if self.list_rounds > 0 and nq > 0:
                # every rank must run the same number of all-gathers: decided from the LONGEST shard's launch plan
                rounds = min(self.list_rounds, self.index.num_launches(nq, k, n_docs=self._max_docs) - 1)
                if rounds > 0:
                    return self.index.topk(q_indptr, q_terms, k, check_status=False, list_exchange=self._gather_running,
                                           list_rounds=rounds)
            return self.index.topk(q_indptr, q_terms, k, check_status=False)
        rounds = self.index.num_launches(q_indptr.numel() - 1, k, n_docs=self._max_docs) - 1
        return self.index.topk(q_indptr, q_terms, k, check_status=False,
                               exchange=lambda theta: exchange_thresholds(theta, self.group), exchange_rounds=rounds)

    def _gather_running(self, run_s: torch.Tensor) -> torch.Tensor:
        """[B,k] running scores of this shard -> [G,B,k] of all shards (one all-gather, buffer reused)."""
        world = _world(self.group)
        key = tuple(run_s.shape)
        buf = self._run_gath.get(key)
        if buf is None:
            if len(self._run_gath) > 16:
                self._run_gath.clear()
            buf = self._run_gath[key] = torch.empty((world,) + key, dtype=torch.float32, device=run_s.device)
        dist.all_gather_into_tensor(buf.view(world * key[0], key[1]), run_s, group=self.group)
        return buf

    def topk(self, q_indptr, q_terms, k: int, check_status: bool = True):
        s, d = self._local_topk(q_indptr, q_terms, k)
        key = (tuple(s.shape), s.device)
        gs, gd = gather_lists(s, d, self.group, self._gath.get(key))
        if len(self._gath) > 64:
            self._gath.clear()
        self._gath[key] = (gs, gd)
        if self._merge is not None:
            return self._merge(gs, gd)
        from .index import merge_topk
        out = merge_topk(gs, gd)
        if check_status and self.index is not None and s.shape[0]:
            self.index.check_status(s.shape[0], k)      # bad term ids raise on the sharded path too
        return out

    def topk_host(self, q_indptr, q_terms, k: int):
        """Host CSR batch -> host (scores, ids, h2d_bytes, d2h_bytes): pinned H2D of the queries, local
        scoring, all-gather and merge ON THE DEVICE, one pinned D2H of the merged lists."""
        if self.index is None:
            raise ValueError("topk_host needs the shard's BM25Index")
        d_qi, d_qt, n_terms = self.index._stage_queries(q_indptr, q_terms)
        s, d = self.topk(d_qi, d_qt, k, check_status=False)
        h_s, h_d = self.index._unstage_lists(s, d)
        self.index.check_status(len(q_indptr) - 1, k)   # synchronises the stream: the copies have landed
        return h_s.numpy().copy(), h_d.numpy().copy(), len(q_indptr) * 8 + n_terms * 4, (len(q_indptr) - 1) * k * 8
