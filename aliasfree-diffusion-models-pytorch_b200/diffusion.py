"""DDPM process of the reference (modules/ddpm_models.py:301-436) with the sampler kept on
the device.

``Diffusion`` has the reference's constructor and methods (noise_images, sample_timesteps,
sample, revert, rotate_2d_matrix, sample_shift).  Differences, all invisible in the results:
  * the Config-E rotation (``theta``) runs as a CUDA kernel instead of
    device -> host -> scipy.ndimage.rotate -> device on each of the 999 steps;
  * the posterior update of Algorithm 1 (:374) is one fused kernel; its three scalar
    coefficients come from a host-side table computed with the reference's fp32 op order;
  * no per-step host tensors are created (the reference builds ``t`` on the CPU each step, :362).
"""
import logging
import weakref

import numpy as np
import torch

from . import ops


class Diffusion:
    def __init__(self, noise_steps=1000, beta_start=1e-4, beta_end=0.02, img_size=256, device="cuda"):
        self.noise_steps, self.beta_start, self.beta_end = noise_steps, beta_start, beta_end
        self.img_size, self.device = img_size, device
        self.beta = self.prepare_noise_schedule().to(device)
        self.alpha = 1.0 - self.beta
        self.alpha_hat = torch.cumprod(self.alpha, dim=0)
        self.filter = None
        # host tables of the per-step scalars of ddpm_models.py:374, same fp32 op order
        a, ah, b = self.alpha.cpu(), self.alpha_hat.cpu(), self.beta.cpu()
        self._ca = (1 / torch.sqrt(a)).tolist()
        self._cb = ((1 - a) / torch.sqrt(1 - ah)).tolist()
        self._cc = torch.sqrt(b).tolist()

    def prepare_noise_schedule(self):
        return torch.linspace(self.beta_start, self.beta_end, self.noise_steps)

    def noise_images(self, x, t, generator=None, noise=None):
        """q-sample (ddpm_models.py:317-321).  ``noise`` injects the draw (tests, exact replays)."""
        sa = torch.sqrt(self.alpha_hat[t])[:, None, None, None]
        sb = torch.sqrt(1 - self.alpha_hat[t])[:, None, None, None]
        if noise is not None:
            eps = noise
        elif generator is not None:
            eps = torch.randn(x.shape, dtype=x.dtype, device=x.device, generator=generator)
        else:
            eps = torch.randn_like(x)
        return sa * x + sb * eps, eps

    def sample_timesteps(self, n, generator=None):
        return torch.randint(low=1, high=self.noise_steps, size=(n,), generator=generator)

    # -- Algorithm 1 --------------------------------------------------------------------
    def _initial(self, n, image_channels, x_init, generator):
        if x_init is not None:
            return x_init.to(self.device, torch.float32).contiguous().clone()
        shape = (n, image_channels, self.img_size, self.img_size)
        # the reference draws the start noise on the CPU generator, then moves it (:360)
        return torch.randn(shape, generator=generator).to(self.device).contiguous()

    def _reverse_step(self, model, x, i, noise):
        t = torch.full((x.shape[0],), i, dtype=torch.long, device=x.device)
        eps = model(x, t).float().contiguous()
        ops.ddpm_update_(x, eps, noise, self._ca[i], self._cb[i], self._cc[i])
        return x

    def _loop(self, model, n, image_channels, theta, keep, x_init, generator, noise_source, progress):
        theta_step = None if theta is None else theta / self.noise_steps
        x = self._initial(n, image_channels, x_init, generator)
        kept = []
        steps = reversed(range(1, self.noise_steps))
        if progress:
            from tqdm import tqdm
            steps = tqdm(steps, position=0, total=self.noise_steps - 1)
        for i in steps:
            noise = None
            if i > 1:
                noise = noise_source(x) if noise_source is not None else torch.randn_like(x)
            x = self._reverse_step(model, x, i, noise)
            if theta_step is not None:
                x = ops.rotate(x, theta_step)
            if keep and i % 100 == 0:
                kept.append(x.clone())
        return x, kept

    # -- one reverse step captured in a CUDA graph --------------------------------------------
    def capture_reverse_step(self, model, x, theta=None):
        """Capture ONE reverse step acting in place on ``x`` (UNet forward, noise draw, posterior
        update, optional rotation, timestep decrement) in a CUDA graph.  Returns ``(graph, step)``:
        ``step`` is the int32 device scalar holding the current timestep (set it to T-1 before the
        first replay); each ``graph.replay()`` advances x by one step with no host work.  The
        (ca, cb, cc) coefficients are read on the device from a [T, 3] table.  Noise comes from
        torch's graph-safe CUDA generator: results match the eager sampler in distribution, not
        bit for bit."""
        dev, n, T = x.device, x.shape[0], self.noise_steps
        cc = list(self._cc)
        if T > 1:
            cc[1] = 0.0                                               # no noise on the final step (:372)
        table = torch.tensor([[a, b, c] for a, b, c in zip(self._ca, self._cb, cc)],
                             dtype=torch.float32, device=dev).contiguous()
        step = torch.full((1,), T - 1, dtype=torch.int32, device=dev)
        t_long = torch.empty(n, dtype=torch.long, device=dev)
        theta_step = None if theta is None else theta / T

        def body():
            t_long.copy_(step.expand(n))
            eps = model(x, t_long).float().contiguous()
            noise = torch.randn_like(x)
            ops.ddpm_update_table_(x, eps, noise, table, step)
            if theta_step is not None:
                x.copy_(ops.rotate(x, theta_step))
            step.sub_(1)

        x_keep = x.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():               # warm-up outside capture (lazy inits)
            for _ in range(2):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        x.copy_(x_keep); step.fill_(T - 1)
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            body()
        x.copy_(x_keep); step.fill_(T - 1)                            # capture itself executes nothing
        graph._afr_keepalive = (table, t_long)                        # tensors the graph reads
        return graph, step

    def prepare_graph(self, model, n, image_channels, theta=None):
        """Capture (once) and cache the reverse-step graph for this (model, batch shape, theta): later
        ``sample(..., cuda_graph=True)`` calls with the same arguments replay it without re-capturing.
        The graph reads the model's parameters in place, so training between calls is fine; a model
        whose parameter STORAGE is replaced (``.to()``, ``load_state_dict`` keeps storage and is fine)
        needs ``clear_graphs()``."""
        key = (id(model), int(n), int(image_channels), None if theta is None else float(theta))
        cache = self.__dict__.setdefault("_graphs", {})
        hit = cache.get(key)
        if hit is None or hit[3]() is not model:            # id() of a collected model may be reused
            x = torch.zeros((n, image_channels, self.img_size, self.img_size), device=self.device)
            graph, step = self.capture_reverse_step(model, x, theta)
            hit = cache[key] = (graph, step, x, weakref.ref(model))
        return hit[:3]

    def clear_graphs(self):
        self.__dict__.pop("_graphs", None)

    def _graphed_loop(self, model, n, image_channels, theta, keep, x_init, generator):
        graph, step, x = self.prepare_graph(model, n, image_channels, theta)
        x.copy_(self._initial(n, image_channels, x_init, generator))
        step.fill_(self.noise_steps - 1)
        kept = []
        for i in reversed(range(1, self.noise_steps)):
            graph.replay()
            if keep and i % 100 == 0:
                kept.append(x.clone())
        return x.clone(), kept

    @staticmethod
    def _to_u8(x):
        return (((x.clamp(-1, 1) + 1) / 2) * 255).type(torch.uint8)

    def sample(self, model, n, image_channels, theta=None, x_init=None, generator=None,
               noise_source=None, progress=False, return_float=False, cuda_graph=False):
        """Ancestral sampling, ddpm_models.py:352-386.  Returns ``(x_u8, result_u8)`` like the
        reference (``result`` = snapshots every 100 steps + the final x).  ``noise_source(x)``
        lets a caller inject the per-step noise (tests replay the reference's CPU stream;
        the sharded sampler slices a full-batch stream).  ``cuda_graph=True`` captures one reverse
        step in a CUDA graph and replays it (launch-bound small batches: several times faster)."""
        logging.info(f"Sampling {n} new images....")
        model.eval()
        with torch.no_grad():
            if cuda_graph:
                if noise_source is not None:
                    raise ValueError("cuda_graph=True draws its noise inside the graph; noise_source is unsupported")
                x, kept = self._graphed_loop(model, n, image_channels, theta, True, x_init, generator)
            else:
                x, kept = self._loop(model, n, image_channels, theta, True, x_init, generator,
                                     noise_source, progress)
        model.train()                       # the reference unconditionally does (:379)
        kept.append(x)
        if return_float:
            return x, torch.cat(kept)
        return self._to_u8(x), self._to_u8(torch.cat(kept))

    def revert(self, model, n, image_channels, **kw):
        """ddpm_models.py:326-350: like ``sample`` without rotation, returns only the snapshots."""
        return self.sample(model, n, image_channels, theta=None, **kw)[1]

    def sample_shift(self, model, n, image_channels, shift=None, **kw):
        """Translation variant (ddpm_models.py:388-419, 'under development' upstream): a
        1-pixel horizontal grid-wrap roll at evenly spaced steps -- ``torch.roll`` on device."""
        if shift == 0:
            shift = None
        idx = set()
        if shift is not None:
            dur = np.abs(shift) / self.noise_steps
            idx = set(np.round(np.arange(0, self.noise_steps, dur)).astype(int)[1:].tolist())
        model.eval()
        with torch.no_grad():
            x = self._initial(n, image_channels, kw.get("x_init"), kw.get("generator"))
            for i in reversed(range(1, self.noise_steps)):
                noise = torch.randn_like(x) if i > 1 else None
                x = self._reverse_step(model, x, i, noise)
                if shift is not None and i in idx:
                    x = torch.roll(x, int(np.sign(shift)), dims=3)
        model.train()
        return self._to_u8(x)

    @staticmethod
    def rotate_2d_matrix(matrix, degrees, filter=None):
        """ddpm_models.py:421-429, on device."""
        return ops.rotate(matrix, degrees)

    @staticmethod
    def shift_2d_matrix(matrix, hshift, vshift, device=None):
        """ddpm_models.py:431-436 (integer grid-wrap shift)."""
        return torch.roll(matrix, (int(vshift), int(hshift)), dims=(2, 3))
