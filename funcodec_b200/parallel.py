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
    is the per-GPU hot path (B200Encodec on the GPU box; any callable in the CPU tests).

    Equal shards (n_clips % world == 0, e.g. BASELINE config 5: 512 clips over 8 GPUs) move without staging copies: the
    scatter reads views of the source batch, the gathers write views of the result tensors; ragged splits are padded to
    the largest shard."""

    def __init__(self, run: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], group=None, src: int = 0):
        self.run = run
        self.group = group
        self.src = src
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._buf = {}

    def _cached(self, key, shape, dtype, device):
        t = self._buf.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != torch.device(device):
            t = torch.empty(shape, dtype=dtype, device=device)
            self._buf[key] = t
        return t

    def __call__(self, wav_all: Optional[torch.Tensor], n_clips: int, length: int, device) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        """wav_all: [n_clips, L] on `src` (device tensor), None elsewhere.  Returns (codes [n_q, n_clips, T'],
        recon [n_clips, 1, L]) on `src`, None elsewhere."""
        bounds = shard_bounds(n_clips, self.world)
        lo, hi = bounds[self.rank]
        maxn = max(b - a for a, b in bounds)
        even = all(b - a == maxn for a, b in bounds)
        is_src = self.rank == self.src
        buf = self._cached("in", (maxn, length), torch.float32, device)
        if is_src:
            if even:
                chunks = [wav_all[a:b] for a, b in bounds]              # contiguous row ranges: no staging copy
            else:
                chunks = []
                for a, b in bounds:
                    c = torch.zeros((maxn, length), dtype=torch.float32, device=device)
                    c[: b - a] = wav_all[a:b]
                    chunks.append(c)
            dist.scatter(buf, chunks, src=self.src, group=self.group)
        else:
            dist.scatter(buf, None, src=self.src, group=self.group)
        mine = buf[: hi - lo]
        if hi > lo:
            codes, recon = self.run(mine)
            n_q, _, tf = codes.shape
        else:
            codes, recon, n_q, tf = None, None, 0, 0
        if even and hi > lo:
            cpad, rpad = codes.contiguous(), recon.contiguous()
        else:
            meta = torch.tensor([n_q, tf], dtype=torch.int64, device=device)
            dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=self.group)
            n_q, tf = int(meta[0]), int(meta[1])
            cpad = torch.zeros((n_q, maxn, tf), dtype=torch.int64, device=device)
            rpad = torch.zeros((maxn, 1, length), dtype=torch.float32, device=device)
            if hi > lo:
                cpad[:, : hi - lo] = codes
                rpad[: hi - lo] = recon
        if is_src:
            call = self._cached("codes", (self.world, n_q, maxn, tf), torch.int64, device)
            rall = self._cached("recon", (self.world * maxn, 1, length), torch.float32, device)
            dist.gather(cpad, [call[r] for r in range(self.world)], dst=self.src, group=self.group)
            dist.gather(rpad, [rall[r * maxn:(r + 1) * maxn] for r in range(self.world)], dst=self.src, group=self.group)
            if even:
                # [world][n_q][b][T'] -> [n_q][world * b][T'] (B is the middle dimension of the reference layout)
                return call.permute(1, 0, 2, 3).reshape(n_q, n_clips, tf), rall
            codes_all = torch.cat([call[r][:, : b - a] for r, (a, b) in enumerate(bounds)], dim=1)
            recon_all = torch.cat([rall[r * maxn: r * maxn + (b - a)] for r, (a, b) in enumerate(bounds)], dim=0)
            return codes_all, recon_all
        dist.gather(cpad, None, dst=self.src, group=self.group)
        dist.gather(rpad, None, dst=self.src, group=self.group)
        return None


class SharedHostBatch:
    """ONE host-resident batch shared by the ranks of a box without funnelling it through one GPU's PCIe link: the batch
    lives in a POSIX shared-memory file (`/dev/shm`), every rank maps it and page-locks its mapping (cudaHostRegister), so
    each GPU DMA-pulls ITS shard of the clips and pushes ITS codes / waveforms over its own PCIe link.  No data-path
    collective at all (a barrier orders producer and consumers); NCCL scatter / gather (ShardedCodec) remains the variant for a
    batch that already lives in one GPU's memory."""

    def __init__(self, name: str, n_clips: int, length: int, n_q: int, frames: int, rank: int, world: int, create: bool):
        import os
        self.n_clips, self.length, self.n_q, self.frames = n_clips, length, n_q, frames
        self.rank, self.world = rank, world
        self.path = os.path.join("/dev/shm", name)
        sizes = [n_clips * length * 4, n_q * n_clips * frames * 8, n_clips * length * 4]
        self.nbytes = sum(sizes)
        if create:
            # reserve the pages now: a tmpfs that is too small must fail HERE (OSError) and not with SIGBUS at the first write
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            try:
                os.posix_fallocate(fd, 0, self.nbytes)
            except OSError:
                os.close(fd)
                try:
                    os.unlink(self.path)
                except OSError:
                    pass
                raise
            os.close(fd)
        self._create = create
        self._raw = None
        self._sizes = sizes

    def map(self):
        """After every rank has seen the file (barrier between create and map)."""
        raw = torch.from_file(self.path, shared=True, size=self.nbytes, dtype=torch.uint8)
        self._raw = raw
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(raw.data_ptr(), self.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed ({rc})")
        o1, o2 = self._sizes[0], self._sizes[0] + self._sizes[1]
        self.wav = raw[:o1].view(torch.float32).view(self.n_clips, self.length)
        # per-rank code slabs [world][n_q][b][T'] so that each rank's D2H destination is contiguous
        b = self.n_clips // self.world
        self.codes = raw[o1:o2].view(torch.int64).view(self.world, self.n_q, b, self.frames)
        self.recon = raw[o2:].view(torch.float32).view(self.n_clips, 1, self.length)
        return self

    def shard(self):
        lo, hi = shard_bounds(self.n_clips, self.world)[self.rank]
        return lo, hi

    def close(self):
        import os
        if self._raw is not None and torch.cuda.is_available():
            torch.cuda.cudart().cudaHostUnregister(self._raw.data_ptr())
        self._raw = None
        if self._create:
            try:
                os.unlink(self.path)
            except OSError:
                pass
