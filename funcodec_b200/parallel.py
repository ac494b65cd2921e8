"""Multi-GPU sharding of independent clips: one process per GPU, torch.distributed for plumbing.

The reference parallelises inference by splitting wav.scp into N contiguous shards and running N independent
processes whose outputs are concatenated (egs/LibriTTS/codec/encoding_decoding.sh:69-100,
funcodec/bin/codec_inference.py:569-579).  The path has NO exchange step (RMS scale, GroupNorm(1,C), LSTM state and
RVQ are all per clip), so the only communication is the scatter of clips from the rank that holds them and the
gather of codes / waveforms back, in the reference's layouts (codes [n_q, B, T'] with B in the middle).
"""
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_clips: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous split like utils/split_scp.pl: the first (n % world) shards get one extra clip."""
    base, extra = divmod(n_clips, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


class ShardedCodec:
    """scatter -> per-rank encode/decode -> gather.  `run(wav[b, L]) -> (codes [n_q, b, T'] int64, recon [b, 1, L])`
    is the per-GPU hot path (B200Encodec on the GPU box; any callable in the CPU tests)."""

    def __init__(self, run: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], group=None, src: int = 0):
        self.run = run
        self.group = group
        self.src = src
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def __call__(self, wav_all: Optional[torch.Tensor], n_clips: int, length: int, device) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        """wav_all: [n_clips, L] on `src` (device tensor), None elsewhere.  Returns (codes [n_q, n_clips, T'],
        recon [n_clips, 1, L]) on `src`, None elsewhere."""
        bounds = shard_bounds(n_clips, self.world)
        lo, hi = bounds[self.rank]
        mine = torch.empty((hi - lo, length), dtype=torch.float32, device=device)
        # scatter requires equal sizes: pad every shard to the largest, trim after
        maxn = max(b - a for a, b in bounds)
        buf = torch.zeros((maxn, length), dtype=torch.float32, device=device)
        if self.rank == self.src:
            chunks = []
            for a, b in bounds:
                c = torch.zeros((maxn, length), dtype=torch.float32, device=device)
                c[: b - a] = wav_all[a:b]
                chunks.append(c)
            dist.scatter(buf, chunks, src=self.src, group=self.group)
        else:
            dist.scatter(buf, None, src=self.src, group=self.group)
        mine.copy_(buf[: hi - lo])
        if hi > lo:
            codes, recon = self.run(mine)
            n_q, _, tf = codes.shape
        else:
            codes, recon, n_q, tf = None, None, 0, 0
        meta = torch.tensor([n_q, tf], dtype=torch.int64, device=device)
        dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=self.group)
        n_q, tf = int(meta[0]), int(meta[1])
        cpad = torch.zeros((n_q, maxn, tf), dtype=torch.int64, device=device)
        rpad = torch.zeros((maxn, 1, length), dtype=torch.float32, device=device)
        if hi > lo:
            cpad[:, : hi - lo] = codes
            rpad[: hi - lo] = recon
        if self.rank == self.src:
            cl = [torch.empty_like(cpad) for _ in range(self.world)]
            rl = [torch.empty_like(rpad) for _ in range(self.world)]
            dist.gather(cpad, cl, dst=self.src, group=self.group)
            dist.gather(rpad, rl, dst=self.src, group=self.group)
            # place each rank's [n_q, b_r, T'] slab at its clip offset (B is the middle dimension)
            codes_all = torch.cat([c[:, : b - a] for c, (a, b) in zip(cl, bounds)], dim=1)
            recon_all = torch.cat([r[: b - a] for r, (a, b) in zip(rl, bounds)], dim=0)
            return codes_all, recon_all
        dist.gather(cpad, None, dst=self.src, group=self.group)
        dist.gather(rpad, None, dst=self.src, group=self.group)
        return None
