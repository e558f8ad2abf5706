"""Row-sharded synchronous SCA over the GPUs of one box (BASELINE config 5).

The MultiSpinFlip SCA of an N-spin model is the bipartite SCA (reference: src/OnBipartiteGraph.jl:30-43) on the
embedding W = (J + qI)/2, sigma = tau = s (demo.jl:82-90).  When W does not fit one GPU, rank g owns the output
units [g*nb, (g+1)*nb) of both half-steps and W[block rows, :]; after every half-step the freshly sampled
[R][nb] blocks are all-gathered (NCCL over NVLink) into the block-major [G][R][nb] matrix that is the next
half-step's K operand.  torch is plumbing here: device buffers, the stream, and the all-gather.
The library's noise is indexed by global (replica, step, unit), so the result does not depend on G: with
``emulate_blocks=G`` one process plays all G ranks in turn on one GPU (used by the parity tests).
"""
from __future__ import annotations

import numpy as np

from . import _lib


class RowShardedSCA:
    def __init__(self, n: int, R: int, *, seed: int | None = None, q: float = 1.0, W=None, h=None, rule=_lib.BIP_SCA,
                 prec=_lib.PREC_BF16X3, emulate_blocks: int | None = None, device: int | None = None, group=None,
                 fused: bool = True):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.n, self.R, self.rule = int(n), int(R), rule
        self.distributed = emulate_blocks is None and dist.is_available() and dist.is_initialized() \
            and dist.get_world_size(group) > 1
        if emulate_blocks is not None:
            self.G, self.blocks = int(emulate_blocks), list(range(int(emulate_blocks)))
        elif self.distributed:
            self.G, self.blocks = dist.get_world_size(group), [dist.get_rank(group)]
        else:
            self.G, self.blocks = 1, [0]
        if self.n % self.G:
            raise ValueError("n must be divisible by the number of blocks")
        self.nb = self.n // self.G
        self.ctx = _lib.context(device)
        self.dev = torch.device("cuda", self.ctx.device)
        torch.cuda.set_device(self.dev)
        self.ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        self.models = []
        for g in self.blocks:
            if W is not None:
                Wg = np.asarray(W, dtype=np.float64)[g * self.nb:(g + 1) * self.nb, :]
                hb = None if h is None else 0.5 * np.asarray(h, dtype=np.float64)[g * self.nb:(g + 1) * self.nb]
                self.models.append(_lib.Model.shard_rows(self.ctx, self.n, self.G, g, Wg, hb, hb, prec))
            else:
                self.models.append(_lib.Model.shard_sk(self.ctx, self.n, self.G, g, int(seed), q, prec))
        bf = torch.bfloat16
        # Fused exchange: the gathered matrices live in symmetric (peer-mapped) memory and every rank's sampling
        # epilogue stores its block straight into all peers' copies over NVLink; a cross-GPU barrier then replaces
        # the all-gather.  Falls back to ncclAllGather when symmetric memory is unavailable.
        self.fused = False
        self.full_v = self.full_h = None
        if self.distributed and fused and 2 <= self.G <= 8:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                self.full_v = symm_mem.empty((self.G, self.R, self.nb), dtype=bf, device=self.dev)
                self.full_h = symm_mem.empty((self.G, self.R, self.nb), dtype=bf, device=self.dev)
                self._hv = symm_mem.rendezvous(self.full_v, grp)
                self._hh = symm_mem.rendezvous(self.full_h, grp)
                self.full_v.zero_()
                self.full_h.zero_()
                self.fused = True
            except Exception as exc:  # pragma: no cover - depends on the driver / torch build
                self.fused_error = repr(exc)
                self.full_v = self.full_h = None
        if self.full_v is None:
            self.full_v = torch.zeros((self.G, self.R, self.nb), dtype=bf, device=self.dev)
            self.full_h = torch.zeros((self.G, self.R, self.nb), dtype=bf, device=self.dev)
        # this rank's freshly sampled blocks, one per layer (they also hold the block's previous values)
        self.blk_v = [torch.ones((self.R, self.nb), dtype=bf, device=self.dev) for _ in self.blocks]
        self.blk_h = [torch.ones((self.R, self.nb), dtype=bf, device=self.dev) for _ in self.blocks]
        self.launches = 0
        self.gather_bytes = 0

    # ---- state
    def set_spins(self, S):
        """S: (R, n) int8 +-1; the embedding starts from sigma = tau = s."""
        torch = self.torch
        S = torch.as_tensor(np.ascontiguousarray(S, dtype=np.int8), device=self.dev)
        full = S.view(self.R, self.G, self.nb).permute(1, 0, 2).contiguous()
        self.full_v.copy_(full.to(torch.bfloat16))
        self.full_h.copy_(self.full_v)
        for i, g in enumerate(self.blocks):
            self.blk_v[i].copy_(self.full_v[g])
            self.blk_h[i].copy_(self.full_v[g])
        if self.fused:
            # no peer may store into this rank's matrices before they hold the initial configuration
            torch.cuda.synchronize(self.dev)
            self.dist.barrier(group=self.group)

    def get_spins(self):
        """(R, n) int8 visible layer (identical on every rank after the all-gather)."""
        s = (self.full_v > 0).to(self.torch.int8) * 2 - 1
        return s.permute(1, 0, 2).reshape(self.R, self.n).cpu().numpy()

    def get_hidden(self):
        s = (self.full_h > 0).to(self.torch.int8) * 2 - 1
        return s.permute(1, 0, 2).reshape(self.R, self.n).cpu().numpy()

    # ---- one half-step: every owned block samples its units, then the blocks are exchanged
    def _half(self, layer, seed, step_abs, T):
        src, dst = (self.full_v, self.full_h) if layer == 1 else (self.full_h, self.full_v)
        if self.fused:
            g, hdl = self.blocks[0], (self._hh if layer == 1 else self._hv)
            slab = g * self.R * self.nb * 2  # byte offset of this rank's block in every gathered matrix
            peers = [int(ptr) + slab for q, ptr in enumerate(hdl.buffer_ptrs) if q != hdl.rank]
            self.models[0].shard_halfstep_fused(self.R, layer, self.rule, src.data_ptr(), dst.data_ptr() + slab, peers, seed,
                                                step_abs, T)
            self.launches += 1
            self.gather_bytes += (self.G - 1) * self.R * self.nb * 2
            hdl.barrier(channel=0)  # all peers' stores have landed before anyone reads the layer
            return
        blk = self.blk_h if layer == 1 else self.blk_v
        for i, m in enumerate(self.models):
            m.shard_halfstep(self.R, layer, self.rule, src.data_ptr(), blk[i].data_ptr(), seed, step_abs, T)
            self.launches += 1
        if self.distributed:
            # the one real exchange step of this path: [R][nb] per rank -> [G][R][nb] everywhere
            self.dist.all_gather_into_tensor(dst.view(-1), blk[0].view(-1), group=self.group)
            self.gather_bytes += (self.G - 1) * blk[0].numel() * 2
        else:
            for i, g in enumerate(self.blocks):
                dst[g].copy_(blk[i])

    def run(self, nsteps, T, *, seed=0, step_offset=0):
        """nsteps synchronous SCA steps; T: array of nsteps temperatures (T[k] applies to step k)."""
        T = np.atleast_1d(np.asarray(T, dtype=np.float64))
        for k in range(int(nsteps)):
            self._half(1, seed, step_offset + k, float(T[min(k, len(T) - 1)]))
            self._half(0, seed, step_offset + k, float(T[min(k, len(T) - 1)]))
