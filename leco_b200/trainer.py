"""The LECO training iteration (train_lora.py:141-302) as a fused, graph-captured GPU step.

`LecoTrainer.iteration()` performs exactly one pass of the reference's loop body:

  RNG draws in the reference's order          train_lora.py:149-156, 175-177
  k-step DDIM denoise under CFG 3 with LoRA on            train_lora.py:179-193, train_util.py:172-193
  timestep remap t* = timesteps_1000[int(k*1000/max)]     train_lora.py:195-199
  positive / neutral / unconditional predictions, LoRA off  train_lora.py:202-237
  target prediction, LoRA on, with gradient               train_lora.py:244-256
  erase / enhance MSE objective                           prompt_util.py:107-135, train_lora.py:265-270
  backward into the LoRA matrices only, AdamW step        train_lora.py:279-281

What it does NOT repeat is work the reference executes redundantly (SURVEY §8a notes):
  * all four predict_noise calls use guidance_scale=1, i.e. e_u + 1*(e_c - e_u) == e_c up to
    rounding, and d/d(e_u) = 0: only the conditional half is evaluated / differentiated;
  * identical prompts among positive / neutral / unconditional are evaluated once;
  * the three LoRA-off passes share (latents, t) and run as one batched forward.
Every kernel launch of a denoise step (UNet forward + CFG/DDIM update) is captured in one CUDA
graph that is replayed k times; the tail (both passes, loss, backward) is a second graph.  The
loss never leaves the device unless the caller reads it.

SDXL (train_lora_xl.py:160-366, train_util.py:217-330): a prompt is an `EmbedsXL` (text [1,77,2048] + pooled
[1,1280]); every UNet call additionally gets `added_cond_kwargs = {text_embeds, time_ids}` where
time_ids = get_add_time_ids(height, width, dynamic_crops) is drawn AFTER the latent noise (:196-201).

Data parallel: every rank draws the SAME global noise / prompt / k from an
identically seeded CPU generator and keeps its slice of the batch; one all-reduce of the
flat fp32 LoRA gradient buffer per iteration, then the same fused AdamW on every rank.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import capi, ops
from .lora import FlatOptimizer, LoRANetwork
from .scheduler import ROW, DDIMScheduler
from .train_util import get_add_time_ids
from .unet import EngineUNet, Tape

UNET_IN_CHANNELS = 4      # train_util.py:12
VAE_SCALE_FACTOR = 8      # train_util.py:13


@dataclass
class EmbedsXL:
    """prompt_util.PromptEmbedsXL (prompt_util.py:17-23): text_embeds [1,77,2048] + pooled_embeds [1,1280]."""
    text_embeds: torch.Tensor
    pooled_embeds: torch.Tensor


@dataclass
class PromptPair:
    """prompt_util.PromptEmbedsPair (prompt_util.py:70-105): four prompt embeddings + settings.  SD1.x/2.x: each
    is a [1,77,D] tensor; SDXL: each is an `EmbedsXL`."""
    target: object
    positive: object
    unconditional: object
    neutral: object
    guidance_scale: float = 1.0
    resolution: int = 512
    dynamic_resolution: bool = False
    batch_size: int = 1
    action: str = "erase"
    dynamic_crops: bool = False     # SDXL only (prompt_util.py:42, train_util.py:301-310)

    def signed_guidance(self) -> float:
        if self.action == "erase":      # neutral - g (positive - unconditional)
            return -float(self.guidance_scale)
        if self.action == "enhance":    # neutral + g (positive - unconditional)
            return float(self.guidance_scale)
        raise ValueError("action must be erase or enhance")


def get_random_resolution_in_bucket(bucket_resolution: int = 512):
    """train_util.py:404-416 (randint upper bound is exclusive)."""
    lo, hi = (bucket_resolution // 2) // 64, bucket_resolution // 64
    h = torch.randint(lo, hi, (1,)).item() * 64
    w = torch.randint(lo, hi, (1,)).item() * 64
    return h, w


class _Graphed:
    """A captured CUDA graph plus the static tensors it reads/writes."""

    def __init__(self):
        self.graph = None
        self.launches = 0  # kernels inside one replay
        self.static: Dict[str, torch.Tensor] = {}


class LecoTrainer:
    def __init__(self, unet: EngineUNet, network: LoRANetwork, scheduler: DDIMScheduler,
                 prompt_pairs: Sequence[PromptPair], *, lr: float = 1e-4, optimizer: str = "adamw",
                 optimizer_kwargs: Optional[dict] = None, lr_scheduler: str = "constant", iterations: int = 1000,
                 lr_scheduler_kwargs: Optional[dict] = None,
                 max_denoising_steps: int = 50, denoise_guidance: float = 3.0, device="cuda",
                 rank: int = 0, world_size: int = 1, use_cuda_graphs: bool = True, state_fp32: bool = False,
                 hoist_cross_kv: bool = True):
        if network.flat is None:
            raise RuntimeError("LecoTrainer needs the flat LoRA layout: move the network to CUDA first "
                               "(network.to('cuda', dtype=torch.bfloat16), or torch.float32 for fp32 master adapters)")
        self.unet, self.network, self.scheduler = unet, network, scheduler
        self.pairs = list(prompt_pairs)
        self.max_steps = max_denoising_steps
        self.denoise_guidance = float(denoise_guidance)
        self.device = torch.device(device)
        self.rank, self.world = rank, world_size
        self.use_graphs = use_cuda_graphs
        # optimizer + LR schedule exactly as train_lora.py:80-95 builds them (train_util.get_optimizer / get_lr_scheduler)
        self.optimizer = FlatOptimizer(network.flat, optimizer, lr=lr, state_fp32=state_fp32, **(optimizer_kwargs or {}))
        from .train_util import get_lr_scheduler
        self.lr_scheduler = get_lr_scheduler(lr_scheduler, self.optimizer, max_iterations=iterations, lr_min=lr / 100,
                                             **(lr_scheduler_kwargs or {}))
        self.hoist_cross_kv = hoist_cross_kv
        # DDIM keeps the lean CFG+update kernel; DDPM / LMS / Euler-a (model_util.py:247-274) use the general one
        self.general_sched = not isinstance(scheduler, DDIMScheduler)
        self.scaled_input = float(scheduler.init_noise_sigma) != 1.0      # sigma schedulers scale the UNet input
        self._pool = None           # one CUDA-graph memory pool shared by every captured shape (dynamic_resolution)
        self._ar_events = []        # (start, end) CUDA events around the data-parallel all-reduce
        self.profile_phases = False  # record CUDA events at the phase boundaries of iteration() (bench.py)
        self._phase_events = []
        self.act_dtype = unet._act_dtype
        self.xl = bool(unet.spec.text_time)
        for p in self.pairs:
            if self.xl != isinstance(p.target, EmbedsXL):
                raise ValueError("SDXL UNets take EmbedsXL prompts (text + pooled); SD1.x/2.x take plain tensors")
        for p in self.pairs:
            if p.batch_size % world_size != 0:
                raise ValueError(f"batch_size {p.batch_size} must be divisible by world_size {world_size}")
        self._emb_cache: Dict[int, torch.Tensor] = {}
        self._den: Dict[tuple, _Graphed] = {}
        self._tail: Dict[tuple, _Graphed] = {}
        self._tables = None
        self.h2d_bytes = 0
        self._pinned: Dict[tuple, torch.Tensor] = {}
        self._iter = 0
        self.launches = 0
        self.last = {}

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _text(e) -> torch.Tensor:
        return e.text_embeds if isinstance(e, EmbedsXL) else e

    @staticmethod
    def _same_prompt(a, b) -> bool:
        if a is b:
            return True
        ta, tb = LecoTrainer._text(a), LecoTrainer._text(b)
        if ta.shape != tb.shape or not torch.equal(ta, tb):
            return False
        if isinstance(a, EmbedsXL) != isinstance(b, EmbedsXL):
            return False
        return not isinstance(a, EmbedsXL) or torch.equal(a.pooled_embeds, b.pooled_embeds)

    def _pooled(self, embs: Sequence[EmbedsXL], bl: int) -> torch.Tensor:
        """concat_embeddings order for the pooled embeddings: [P] rows, each prompt repeated bl times."""
        return torch.cat([e.pooled_embeds.to(self.device, dtype=self.act_dtype).reshape(1, -1).expand(bl, -1)
                          for e in embs], 0).contiguous()

    def _added(self, st, n_rows: int, key_pooled: str):
        """added_cond_kwargs of one UNet call over the first n_rows samples (None for SD1.x/2.x)."""
        if not self.xl:
            return None
        return {"text_embeds": st[key_pooled], "time_ids": st["ids"][:n_rows]}

    def _emb(self, e: torch.Tensor) -> torch.Tensor:
        k = id(e)
        if k not in self._emb_cache:
            self._emb_cache[k] = e.to(self.device, dtype=self.act_dtype).reshape(-1, e.shape[-1]).contiguous()
        return self._emb_cache[k]

    def _ctx(self, embs: Sequence[torch.Tensor], bl: int) -> torch.Tensor:
        """train_util.concat_embeddings order: each embedding repeated bl times, blocks concatenated."""
        embs = [self._text(e) for e in embs]
        return torch.cat([self._emb(e).unsqueeze(0).expand(bl, -1, -1) for e in embs], 0).reshape(
            -1, embs[0].shape[-1]).contiguous()

    def _sched_tables(self):
        """Device tables for the 50-step grid: timestep and (guidance, cx, ce) per step."""
        s = self.scheduler
        s.set_timesteps(self.max_steps)
        ts = [float(t) for t in s.timesteps]
        coef = s.table(self.denoise_guidance)       # rows [guidance, cx, ce, cn, in_scale, dx, dg, l0..l3, slot]
        self._tables = (torch.tensor(ts, dtype=torch.float32, device=self.device),
                        torch.tensor(coef, dtype=torch.float32, device=self.device))
        return self._tables

    def _graph_pool(self):
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        return self._pool

    # ------------------------------------------------------------------ one denoise step
    def _denoise_body(self, st):
        x_in = ops.scale_by_dev(st["x"], st["coef"], 4) if self.scaled_input else st["x"]   # scale_model_input
        x2 = x_in.repeat(2, 1, 1, 1)
        eps = self.unet.run(x2, st["t"], st["ctx"], self._added(st, x2.shape[0], "pooled"), None,
                            kv_cache=st.get("kv"))
        if self.general_sched:
            ops.sched_step(eps, st["x"], st["coef"], noise=st.get("noise"), hist=st.get("hist"), out=st["x"])
        else:
            x_new, _ = ops.guided_step(eps, st["x"], st["coef"], True, False)
            st["x"].copy_(x_new)

    def _denoise_graph(self, bl, h, w, D):
        key = (bl, h, w, D)
        g = self._den.get(key)
        if g is not None:
            return g
        g = _Graphed()
        dev = self.device
        g.static = {"x": torch.zeros((bl, UNET_IN_CHANNELS, h, w), device=dev, dtype=torch.float32),
                    "t": torch.zeros((2 * bl,), device=dev, dtype=torch.float32),
                    "coef": torch.zeros((ROW,), device=dev, dtype=torch.float32),
                    "ctx": torch.zeros((2 * bl * 77, D), device=dev, dtype=self.act_dtype)}
        if self.scheduler.needs_noise:
            g.static["noise"] = torch.zeros((bl, UNET_IN_CHANNELS, h, w), device=dev, dtype=torch.float32)
        if self.scheduler.history:
            g.static["hist"] = torch.zeros((4, bl * UNET_IN_CHANNELS * h * w), device=dev, dtype=torch.float32)
        if self.xl:
            g.static["pooled"] = torch.zeros((2 * bl, self.unet.spec.add_text_dim), device=dev, dtype=self.act_dtype)
            g.static["ids"] = torch.zeros((2 * bl, 6), device=dev, dtype=torch.float32)
        if self.hoist_cross_kv:
            # K|V of the text embedding per cross-attention layer: constant over the k steps of one iteration
            # (they depend on the prompt and the adapter state only), computed once by iteration() into these buffers
            g.static["kv"] = [torch.zeros(shape, device=dev, dtype=self.act_dtype)
                              for shape in self.unet.cross_kv_shapes(2 * bl * 77)]
        if self.use_graphs:
            self.network.__enter__()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._denoise_body(g.static)  # warm-up (lazy inits, allocator)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g.graph = torch.cuda.CUDAGraph()
            c0 = capi.launch_count()
            with torch.cuda.graph(g.graph, pool=self._graph_pool()):
                self._denoise_body(g.static)
            g.launches = capi.launch_count() - c0
            self.network.__exit__(None, None, None)
        self._den[key] = g
        return g

    # ------------------------------------------------------------------ tail: 4 predictions + loss + backward
    def _tail_body(self, st, groups, slots, bl, sgn_g):
        """groups = number of distinct LoRA-off prompts, slots = (i_pos, i_neu, i_unc) into them."""
        net = self.network
        net.__exit__(None, None, None)                           # multiplier 0: LoRA-off passes
        x_in = ops.scale_by_dev(st["x"], st["in_scale"], 0) if self.scaled_input else st["x"]   # scale_model_input at t*
        xr = x_in.repeat(groups, 1, 1, 1)
        eps_ng = self.unet.run(xr, st["t"][: groups * bl], st["ctx_ng"], self._added(st, groups * bl, "pooled_ng"), None)
        pos, neu, unc = (eps_ng[i * bl:(i + 1) * bl] for i in slots)
        net.__enter__()                                          # multiplier 1: target pass with tape
        tape = Tape(ops)
        eps_t = self.unet.run(x_in, st["t"][:bl], st["ctx_t"], self._added(st, bl, "pooled_t"), tape)
        loss, dt = ops.leco_loss(eps_t, pos, neu, unc, sgn_g, True)
        tape.grads["eps"] = dt
        tape.backward()
        net.__exit__(None, None, None)
        st["loss"].copy_(loss)
        st["eps_t"].copy_(eps_t)

    def _tail_graph(self, bl, h, w, D, groups, slots, sgn_g):
        key = (bl, h, w, D, groups, slots, sgn_g)
        g = self._tail.get(key)
        if g is not None:
            return g
        g = _Graphed()
        dev = self.device
        g.static = {"x": torch.zeros((bl, UNET_IN_CHANNELS, h, w), device=dev, dtype=torch.float32),
                    "t": torch.zeros((groups * bl,), device=dev, dtype=torch.float32),
                    "ctx_ng": torch.zeros((groups * bl * 77, D), device=dev, dtype=self.act_dtype),
                    "ctx_t": torch.zeros((bl * 77, D), device=dev, dtype=self.act_dtype),
                    "loss": torch.zeros((1,), device=dev, dtype=torch.float32),
                    "eps_t": torch.zeros((bl, UNET_IN_CHANNELS, h, w), device=dev, dtype=torch.float32),
                    "in_scale": torch.ones((1,), device=dev, dtype=torch.float32)}
        if self.xl:
            P = self.unet.spec.add_text_dim
            g.static["pooled_ng"] = torch.zeros((groups * bl, P), device=dev, dtype=self.act_dtype)
            g.static["pooled_t"] = torch.zeros((bl, P), device=dev, dtype=self.act_dtype)
            g.static["ids"] = torch.zeros((groups * bl, 6), device=dev, dtype=torch.float32)
        if self.use_graphs:
            # a new shape can first appear while gradients are being accumulated (step_optimizer=False, or
            # dynamic_resolution mid-run): the warm-up pass must not disturb them
            keep = self.network.flat.grads.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._tail_body(g.static, groups, slots, bl, sgn_g)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.network.flat.grads.copy_(keep)                  # drop what the warm-up accumulated
            del keep
            g.graph = torch.cuda.CUDAGraph()
            c0 = capi.launch_count()
            with torch.cuda.graph(g.graph, pool=self._graph_pool()):
                self._tail_body(g.static, groups, slots, bl, sgn_g)
            g.launches = capi.launch_count() - c0
        self._tail[key] = g
        return g

    # ------------------------------------------------------------------ the iteration
    @torch.no_grad()
    def iteration(self, fixed_k: Optional[int] = None, step_optimizer: bool = True,
                  device_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One LECO iteration; returns the (device) loss.  `device_noise` [B_local,4,h,w] fp32 skips the
        host draw + H2D copy of the latent noise (inputs already resident in HBM)."""
        sched = self.scheduler
        launches0 = capi.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if self.profile_phases else None
        if ev:
            ev[0].record()
        self.network.flat.refresh_transposed()       # ad^T / bup^T for the backward GEMMs (one launch)
        tbl_t, tbl_coef = self._tables or self._sched_tables()
        sched.set_timesteps(self.max_steps)
        # ---- RNG draws: same generator, same order as train_lora.py:149-177
        pair = self.pairs[torch.randint(0, len(self.pairs), (1,)).item()]
        k = torch.randint(1, self.max_steps, (1,)).item()
        if fixed_k is not None:
            k = int(fixed_k)
        height = width = pair.resolution
        if pair.dynamic_resolution:
            height, width = get_random_resolution_in_bucket(pair.resolution)
        h, w = height // VAE_SCALE_FACTOR, width // VAE_SCALE_FACTOR
        bl = pair.batch_size // self.world
        D = self._text(pair.target).shape[-1]
        dg = self._denoise_graph(bl, h, w, D)
        st = dg.static
        if device_noise is not None:
            st["x"].copy_(device_noise)
            self.h2d_bytes = 0
        else:
            noise = torch.randn((pair.batch_size, UNET_IN_CHANNELS, h, w), device="cpu") * sched.init_noise_sigma
            self._stage_noise(noise, bl, h, w, st["x"])
        st["ctx"].copy_(self._ctx([pair.unconditional, pair.target], bl))
        ids = None
        if self.xl:  # drawn after the latent noise (train_lora_xl.py:183-201); same row for every sample
            ids = get_add_time_ids(height, width, dynamic_crops=pair.dynamic_crops).to(self.device, dtype=torch.float32)
            st["ids"].copy_(ids.reshape(1, 6).expand(2 * bl, 6))
            st["pooled"].copy_(self._pooled([pair.unconditional, pair.target], bl))
        if "kv" in st:                               # once per iteration instead of once per denoise step
            self.network.__enter__()
            self.unet.cross_kv(st["ctx"], out=st["kv"])
            self.network.__exit__(None, None, None)
        step_noise = None
        if sched.needs_noise:
            # diffusers draws randn on the model's device every step (global CUDA generator); one draw of the GLOBAL
            # batch for all k steps, sliced per rank, keeps an R-GPU run identical to the 1-GPU run
            step_noise = torch.randn((k, pair.batch_size, UNET_IN_CHANNELS, h, w), device=self.device)[
                :, self.rank * bl:(self.rank + 1) * bl]
        if dg.graph is None:
            self.network.__enter__()
        for i in range(k):
            st["t"].copy_(tbl_t[i].expand(2 * bl))
            st["coef"].copy_(tbl_coef[i])
            if step_noise is not None:
                st["noise"].copy_(step_noise[i])
            if dg.graph is not None:
                dg.graph.replay()
                self.launches += dg.launches
            else:
                self._denoise_body(st)
        if dg.graph is None:
            self.network.__exit__(None, None, None)

        if ev:
            ev[1].record()
        # ---- tail at t* = timesteps_1000[int(k*1000/max)] = 999 - int(k*1000/max)
        t_star = float(999 - int(k * 1000 / self.max_steps))
        distinct: List[torch.Tensor] = []
        slots = []
        for e in (pair.positive, pair.neutral, pair.unconditional):
            for j, d in enumerate(distinct):
                if self._same_prompt(d, e):
                    slots.append(j)
                    break
            else:
                distinct.append(e)
                slots.append(len(distinct) - 1)
        tg = self._tail_graph(bl, h, w, D, len(distinct), tuple(slots), pair.signed_guidance())
        ts = tg.static
        ts["x"].copy_(st["x"])
        ts["t"].fill_(t_star)
        if self.scaled_input:
            ts["in_scale"].fill_(sched.in_scale_at_train_timestep(int(t_star)))
        ts["ctx_ng"].copy_(self._ctx(distinct, bl))
        ts["ctx_t"].copy_(self._ctx([pair.target], bl))
        if self.xl:
            ts["ids"].copy_(ids.reshape(1, 6).expand(ts["ids"].shape[0], 6))
            ts["pooled_ng"].copy_(self._pooled(distinct, bl))
            ts["pooled_t"].copy_(self._pooled([pair.target], bl))
        if tg.graph is not None:
            tg.graph.replay()
            self.launches += tg.launches
        else:
            self._tail_body(ts, len(distinct), tuple(slots), bl, pair.signed_guidance())
        loss = ts["loss"]

        # ---- data-parallel exchange: ONE all-reduce of the flat fp32 LoRA gradient; the scalar loss rides in the
        # buffer's last slot
        if self.world > 1:
            import torch.distributed as dist
            flat = self.network.flat
            flat.loss_slot.copy_(loss)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(flat.grads_ext)
            e1.record()
            self._ar_events = (self._ar_events + [(e0, e1)])[-32:]
            loss = flat.loss_slot * (1.0 / self.world)
        if step_optimizer:
            self.optimizer.step(grad_scale=1.0 / self.world)
            self.lr_scheduler.step()                          # train_lora.py:281
        if ev:
            ev[2].record()
            self._phase_events = (self._phase_events + [(k, ev)])[-64:]
        self.launches += capi.launch_count() - launches0     # eager launches (graph replays were added above)
        self.last = {"k": k, "timestep": t_star, "denoised": st["x"], "pair": pair, "target": ts["eps_t"],
                     "lr": self.optimizer.lr}
        return loss

    def phase_ms(self):
        """Mean device ms of (the k-step denoise loop, everything after it: 4 predictions, loss, backward, optimizer)
        and the mean k over the recorded iterations (profile_phases=True)."""
        if not self._phase_events:
            return None
        torch.cuda.synchronize()
        n = len(self._phase_events)
        den = sum(e[0].elapsed_time(e[1]) for _, e in self._phase_events) / n
        tail = sum(e[1].elapsed_time(e[2]) for _, e in self._phase_events) / n
        return {"denoise_loop_ms": den, "tail_ms": tail, "mean_k": sum(k for k, _ in self._phase_events) / n}

    def allreduce_ms(self) -> Optional[float]:
        """Median device time of the data-parallel all-reduce over the last iterations (None on one rank).  Event to
        event on this rank, so it includes waiting for the slowest rank; the median keeps NCCL's lazy communicator
        set-up inside the very first call out of the figure."""
        if not self._ar_events:
            return None
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in self._ar_events)
        return ts[len(ts) // 2]

    def _stage_noise(self, noise, bl, h, w, dst):
        """Host noise slice of this rank -> pinned staging -> device (the only per-step H2D traffic)."""
        ring = self._pinned.get((bl, h, w))
        if ring is None:  # two pinned staging buffers: the host may run one iteration ahead of the GPU
            ring = self._pinned[(bl, h, w)] = [
                [torch.empty((bl, UNET_IN_CHANNELS, h, w), dtype=torch.float32).pin_memory(), None] for _ in range(2)]
        slot = ring[self._iter & 1]
        if slot[1] is not None:
            slot[1].synchronize()
        slot[0].copy_(noise[self.rank * bl:(self.rank + 1) * bl])
        dst.copy_(slot[0], non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        self._iter += 1
        self.h2d_bytes = slot[0].numel() * 4
