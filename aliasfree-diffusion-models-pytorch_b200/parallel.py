"""Multi-GPU use of the path: one process per GPU, ``torch.distributed`` for plumbing.

* Sampling shards the image batch across ranks with NO communication during the 999-step
  loop (every image is independent: GroupNorm(1, C) and attention are per-sample); an
  optional final ``all_gather`` collects the uint8 images (4096 x 3 x 32 x 32 = 12.6 MB).
* Training is data parallel with ONE NCCL all-reduce per step over a flat gradient buffer
  (5.9 M fp32 = 23.6 MB), issued after backward.  The reference has no distributed code
  (SURVEY.md section 2); both wrappers are new code around its unchanged inner loops
  (modules/ddpm_models.py:359-378 and modules/ddpm_utils.py:498-509).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous, balanced split of ``n`` items: first ``n % world`` ranks get one extra."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedNoise:
    """Per-step noise source that reproduces an UNSHARDED run bit for bit: every rank draws
    the full-batch tensor from an identically seeded generator and keeps its slice
    (50 MB/step at n=4096 -- cheap next to a UNet forward)."""

    def __init__(self, full_shape, lo, hi, device, seed):
        self.shape, self.lo, self.hi = tuple(full_shape), lo, hi
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)
        self.device = device

    def __call__(self, x):
        full = torch.randn(self.shape, device=self.device, generator=self.gen)
        return full[self.lo:self.hi].contiguous()


def sharded_sample(diffusion, model, n, image_channels, theta=None, seed=0, gather=False,
                   exact_stream=True, progress=False, cuda_graph=False, memory_format=None):
    """Batch-sharded Algorithm 1.  Rank r samples images [lo, hi) of the global batch ``n``.
    The start noise is the slice of the seeded CPU draw the unsharded reference would make
    (modules/ddpm_models.py:360).  Returns this rank's ``(x_u8, result_u8)``; with
    ``gather=True`` rank-ordered full tensors on every rank (equal shard sizes required).
    ``cuda_graph=True`` replays one captured reverse step per iteration (per-rank device noise).
    ``memory_format=torch.channels_last`` converts the model in place first: cuDNN then runs its convolutions
    without nchw<->nhwc conversion kernels and the fused activations take the channels-last kernel (-8 % per
    reverse step at 4096 images; results agree to convolution rounding)."""
    if memory_format is not None:
        model = model.to(memory_format=memory_format)
    rank, ws = world()
    lo, hi = shard_bounds(n, rank, ws)
    g = torch.Generator()
    g.manual_seed(seed)
    S = diffusion.img_size
    x0 = torch.randn((n, image_channels, S, S), generator=g)[lo:hi]
    noise = None
    if cuda_graph:
        # the captured step draws its own noise: seed the device generator per rank (distributional,
        # not bitwise, agreement with an unsharded run)
        torch.cuda.manual_seed(seed + 1 + rank)
    elif exact_stream:
        noise = ShardedNoise((n, image_channels, S, S), lo, hi, diffusion.device, seed + 1)
    x, result = diffusion.sample(model, hi - lo, image_channels, theta=theta, x_init=x0,
                                 noise_source=noise, progress=progress, cuda_graph=cuda_graph)
    if gather and ws > 1:
        if n % ws:
            raise ValueError("gather=True needs n divisible by the world size")
        xs = [torch.empty_like(x) for _ in range(ws)]
        dist.all_gather(xs, x)
        x = torch.cat(xs)
        k = result.shape[0] // (hi - lo)
        rs = [torch.empty_like(result) for _ in range(ws)]
        dist.all_gather(rs, result)
        # result is [k snapshots x shard]; re-interleave to [k x n]
        result = torch.cat([torch.cat([r.view(k, hi - lo, *r.shape[1:])[s] for r in rs]) for s in range(k)])
    return x, result


class FlatGradAllReduce:
    """Data-parallel wrapper: parameters broadcast from rank 0 once, then after every
    backward ONE all-reduce (mean) over a single flat fp32 gradient buffer whose views are
    the parameters' ``.grad`` tensors (no per-step packing copies)."""

    def __init__(self, model):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        rank, ws = world()
        self.world_size = ws
        if ws > 1:
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0)
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            chunk = self.flat[off:off + p.numel()]
            # the gradient view takes the parameter's own strides (channels-last conv weights keep a channels-last
            # gradient: optimizers with fused kernels insist on matching layouts)
            dense_cl = p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous()
            p.grad = chunk.as_strided(p.size(), p.stride()) if dense_cl else chunk.view_as(p)
            off += p.numel()

    def zero_grad(self):
        self.flat.zero_()

    def sync(self):
        """Average gradients across ranks: the single collective of a training step."""
        if self.world_size > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(self.world_size)


def train_step(model, diffusion, optimizer, images, ddp=None, generator=None, noise=None):
    """Inner step of the reference ``train`` (modules/ddpm_utils.py:498-509): timesteps,
    q-sample, forward, MSE, backward, [grad all-reduce], AdamW.  Returns the loss tensor
    (no host sync here; the reference's two ``loss.item()`` calls are the caller's business).
    ``noise`` replaces the q-sample draw (exact replays)."""
    t = diffusion.sample_timesteps(images.shape[0], generator=generator).to(images.device)
    x_t, noise = diffusion.noise_images(images, t, noise=noise)
    pred = model(x_t, t)
    loss = torch.nn.functional.mse_loss(pred, noise)
    if ddp is not None:
        ddp.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=False)
    loss.backward()
    if ddp is not None:
        ddp.sync()
    optimizer.step()
    return loss


class GraphedTrainStep:
    """The same inner step with its device work captured in two CUDA graphs -- (q-sample, forward,
    MSE, backward into the flat gradient buffer) and (AdamW) -- replayed every step with the single
    gradient all-reduce in between.  At the per-GPU batch of BASELINE configs[1] (256 / 8 = 32 images)
    the eager step is launch-bound (~19 ms for ~4 ms of device work); replay removes that.

    The optimizer must be built with ``capturable=True``.  Timesteps are still drawn on the host
    (reference: ``sample_timesteps`` is a CPU ``randint``) and copied into a static tensor.

    The eager warm-up iterations that precede capture (lazy cuDNN / cuBLAS initialisation, optimizer
    state allocation) run real optimizer steps on a dummy batch; parameters, buffers and every
    optimizer state tensor are snapshotted before and restored IN PLACE afterwards (the graphs hold
    their addresses), so constructing this object leaves the training state untouched.

    ``static_noise=True`` makes the q-sample noise an input (``__call__(images, noise=...)``) instead
    of a draw inside the graph -- used to replay an eager run exactly."""

    def __init__(self, model, diffusion, optimizer, batch_shape, ddp=None, warmup=3, static_noise=False):
        self.model, self.diffusion, self.opt = model, diffusion, optimizer
        self.ddp = ddp if ddp is not None else FlatGradAllReduce(model)
        dev = next(model.parameters()).device
        self.images = torch.zeros(batch_shape, device=dev)
        self.t = torch.ones(batch_shape[0], dtype=torch.long, device=dev)
        self.loss = torch.zeros((), device=dev)
        self.noise = torch.zeros(batch_shape, device=dev) if static_noise else None

        def fwd_bwd():
            self.ddp.zero_grad()
            x_t, noise = diffusion.noise_images(self.images, self.t, noise=self.noise)
            loss = torch.nn.functional.mse_loss(model(x_t, self.t), noise)
            loss.backward()
            self.loss.copy_(loss.detach())

        tensors = list(model.parameters()) + list(model.buffers())
        saved = [t.detach().clone() for t in tensors]
        saved_opt = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                     for p, st in optimizer.state.items()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # eager warm-up: lazy inits, optimizer state
            for _ in range(warmup):
                fwd_bwd()
                self.ddp.sync()
                optimizer.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.g_fb, self.g_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb):
            fwd_bwd()
        with torch.cuda.graph(self.g_opt, pool=self.g_fb.pool()):
            optimizer.step()
        with torch.no_grad():                               # undo the warm-up (capture executes nothing)
            for t, s in zip(tensors, saved):
                t.copy_(s)
            for p, st in optimizer.state.items():
                old = saved_opt.get(p)
                for k, v in st.items():
                    if not torch.is_tensor(v):
                        if old is not None and k in old:
                            st[k] = old[k]
                    elif old is not None and k in old:
                        v.copy_(old[k])
                    else:
                        v.zero_()                           # state created by the warm-up: a fresh optimizer starts at 0
            self.ddp.zero_grad()

    def __call__(self, images, generator=None, noise=None):
        tl = self.timeline
        if tl is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
        self.images.copy_(images, non_blocking=True)
        self.t.copy_(self.diffusion.sample_timesteps(images.shape[0], generator=generator), non_blocking=True)
        if self.noise is not None:
            if noise is None:
                raise ValueError("static_noise=True: pass the q-sample noise to every call")
            self.noise.copy_(noise, non_blocking=True)
        if tl is not None:
            ev[1].record()
        self.g_fb.replay()
        if tl is not None:
            ev[2].record()
        self.ddp.sync()
        if tl is not None:
            ev[3].record()
        self.g_opt.replay()
        if tl is not None:
            ev[4].record()
            tl.append(ev)
        return self.loss

    timeline = None     # set to [] to collect per-step CUDA events (see ``timeline_ms``)

    def timeline_ms(self):
        """Mean device time per phase over the steps recorded since ``timeline = []`` (synchronises):
        host timesteps + H2D copies | graph 1 (q-sample, forward, MSE, backward) | gradient all-reduce |
        graph 2 (AdamW)."""
        torch.cuda.synchronize()
        names = ("inputs_h2d", "graph_fwd_bwd", "grad_allreduce", "graph_adamw")
        if not self.timeline:
            return {}
        out = {n: sum(e[i].elapsed_time(e[i + 1]) for e in self.timeline) / len(self.timeline)
               for i, n in enumerate(names)}
        out["steps"] = len(self.timeline)
        return out
