"""FlowModel with the reference's API (models_con/flow_model.py:59-374): FlowModel(cfg.model),
.encode(batch), .forward(batch) -> dict of six losses, .sample(batch, num_steps, sample_bb, sample_ang,
sample_seq) -> list of per-step dicts of CPU tensors; identical state_dict keys.

sample(): encode once, then per Euler step ONE composite denoiser call (pf_ga_encoder_forward) plus two
small kernels (pf_denoise_post, pf_euler_step).  The trajectory stays on the device and is copied to pinned
host memory once at the end - the reference's nine `.cpu()` synchronisations per step
(flow_model.py:313-314) are gone, the returned structure is the same.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops, so3_utils, torus
from .constants import BBHeavyAtom, max_num_heavyatoms, torsions_mask
from .edge import EdgeEmbedder
from .ga import GAEncoder
from .geometry import construct_3d_basis
from .layers import categorical_from_uniform, clampped_one_hot, sample_from
from .node import NodeEmbedder


def uniform_so3(num_batch, num_res, device="cpu", generator=None):
    """Haar-uniform rotations (the reference draws them on the host with SciPy, pepflow/modules/so3/dist.py:40-45;
    here: normalised 4-D Gaussians on the device -> quat_to_rot kernel)."""
    q = torch.randn(num_batch, num_res, 4, device=device, generator=generator)
    q = q / torch.linalg.norm(q, dim=-1, keepdim=True)
    return ops.quat_to_rot(q)


class FlowModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self._model_cfg = cfg.encoder
        self._interpolant_cfg = cfg.interpolant
        self.node_embedder = NodeEmbedder(cfg.encoder.node_embed_size, max_num_heavyatoms)
        self.edge_embedder = EdgeEmbedder(cfg.encoder.edge_embed_size, max_num_heavyatoms)
        self.ga_encoder = GAEncoder(cfg.encoder.ipa)
        self.sample_structure = self._interpolant_cfg.sample_structure
        self.sample_sequence = self._interpolant_cfg.sample_sequence
        self.K = self._interpolant_cfg.seqs.num_classes
        self.k = self._interpolant_cfg.seqs.simplex_value
        if self.K != 20:
            raise ValueError("the Euler kernels are specialised to 20 residue classes")
        # constants of the loss path as (non-persistent) buffers: no host->device copies inside forward(), which keeps
        # a training iteration capturable in a CUDA graph; the state_dict is unchanged
        self.register_buffer("_ideal_bb", torch.tensor(FlowModel._IDEAL_BB), persistent=False)
        self.register_buffer("_torsions_mask", torsions_mask.clone(), persistent=False)

    # ------------------------------------------------------------------ reference helpers
    def encode(self, batch, autograd=False):
        """flow_model.py:75-93.  autograd=False: the embedder kernels (sampling; call under torch.no_grad());
        autograd=True: their differentiable torch formulations (FlowModel.forward's gradient path)."""
        pos = batch["pos_heavyatom"]
        rotmats_1 = construct_3d_basis(pos[:, :, BBHeavyAtom.CA], pos[:, :, BBHeavyAtom.C], pos[:, :, BBHeavyAtom.N])
        trans_1 = pos[:, :, BBHeavyAtom.CA]
        seqs_1 = batch["aa"]
        angles_1 = batch["torsion_angle"]
        context_mask = torch.logical_and(batch["mask_heavyatom"][:, :, BBHeavyAtom.CA], ~batch["generate_mask"])
        structure_mask = context_mask if self.sample_structure else None
        sequence_mask = context_mask if self.sample_sequence else None
        args = (batch["aa"], batch["res_nb"], batch["chain_nb"], pos, batch["mask_heavyatom"])
        node_fn = self.node_embedder.forward_autograd if autograd else self.node_embedder
        edge_fn = self.edge_embedder.forward_autograd if autograd else self.edge_embedder
        node_embed = node_fn(*args, structure_mask=structure_mask, sequence_mask=sequence_mask)
        edge_embed = edge_fn(*args, structure_mask=structure_mask, sequence_mask=sequence_mask)
        return rotmats_1, trans_1, angles_1, seqs_1, node_embed, edge_embed

    def zero_center_part(self, pos, gen_mask, res_mask):
        """flow_model.py:95-106.  CUDA tensors take the kernel (pf_zero_center); the torch formula is kept only for the
        CPU run of the training harness (world_size-2 gloo test), which never reaches the sampling path."""
        if pos.is_cuda and not pos.requires_grad:
            return ops.zero_center(pos, gen_mask != 0, res_mask)
        center = torch.sum(pos * gen_mask[..., None], dim=1) / (torch.sum(gen_mask, dim=-1, keepdim=True) + 1e-8)
        center = center.unsqueeze(1)
        pos = (pos - center) * res_mask[..., None]
        return pos, center

    def seq_to_simplex(self, seqs):
        return clampped_one_hot(seqs, self.K).float() * self.k * 2 - self.k

    # ------------------------------------------------------------------ sampling
    def init_noise(self, batch, enc, sample_bb=True, sample_ang=True, sample_seq=True, generator=None):
        """Initial state of the flow (flow_model.py:251-273).  Returns a dict of device tensors."""
        rotmats_1, trans_1, angles_1, seqs_1 = enc[:4]
        gm = batch["generate_mask"]
        B, L = seqs_1.shape
        dev = seqs_1.device
        seqs_1_simplex = self.seq_to_simplex(seqs_1)
        if sample_bb:
            rot0 = torch.where(gm[..., None, None], uniform_so3(B, L, device=dev, generator=generator), rotmats_1)
            tr0 = torch.randn((B, L, 3), device=dev, generator=generator)
            tr0, _ = self.zero_center_part(tr0, gm, batch["res_mask"])
            tr0 = torch.where(gm[..., None], tr0, trans_1)
        else:
            rot0, tr0 = rotmats_1.detach().clone(), trans_1.detach().clone()
        if sample_ang:
            ang0 = torus.tor_random_uniform(angles_1.shape, device=dev, dtype=angles_1.dtype, generator=generator)
            ang0 = torch.where(gm[..., None], ang0, angles_1)
        else:
            ang0 = angles_1.detach().clone()
        if sample_seq:
            sx0 = self.k * torch.randn((B, L, self.K), device=dev, generator=generator)
            s0 = sample_from(F.softmax(sx0, dim=-1), generator=generator)
            s0 = torch.where(gm, s0, seqs_1)
            sx0 = torch.where(gm[..., None], sx0, seqs_1_simplex)
        else:
            s0, sx0 = seqs_1.detach().clone(), seqs_1_simplex.detach().clone()
        return {"rotmats_0": rot0, "trans_0": tr0, "angles_0": ang0, "seqs_0": s0, "seqs_0_simplex": sx0}

    @torch.no_grad()
    def sampler_init(self, batch, num_steps=100, sample_bb=True, sample_ang=True, sample_seq=True, *, noise=None,
                     uniforms=None, seed=None, encoded=None, stream_to_host=False, graph=True):
        """Everything FlowModel.sample does before its loop (flow_model.py:229-285): encode, initial noise,
        time grid, device-resident trajectory buffers.  Returns an EulerSampler whose step(n) is one loop
        iteration - bench.py times exactly that call."""
        dev = batch["aa"].device
        if dev.type != "cuda":
            raise RuntimeError("FlowModel.sample needs the batch on a CUDA device (no CPU fallback)")
        enc = encoded if encoded is not None else self.encode(batch)
        if noise is None:
            noise = self.init_noise(batch, enc, sample_bb, sample_ang, sample_seq)
        if seed is None:
            # the categorical draws follow torch's global generator like the reference's torch.multinomial: a fresh
            # Philox key per call, reproducible under seed_all / torch.manual_seed
            seed = int(torch.randint(0, 2 ** 62, (), dtype=torch.int64))
        return EulerSampler(self, batch, enc, noise, num_steps, (sample_bb, sample_ang, sample_seq), uniforms, seed,
                            stream_to_host=stream_to_host, graph=graph)

    @torch.no_grad()
    def sample(self, batch, num_steps=100, sample_bb=True, sample_ang=True, sample_seq=True, *, noise=None,
               uniforms=None, seed=None, encoded=None, graph=True):
        """Euler sampler (flow_model.py:229-374).  Extra keyword-only hooks (all optional):
        noise: dict from init_noise() to inject the initial state; uniforms: [num_steps, 2, B, L] injected
        U[0,1) for the two categorical draws of each step (else Philox keyed by `seed`; None = a fresh key drawn
        from torch's generator); encoded: output of encode(); graph: replay one captured CUDA graph per iteration."""
        smp = self.sampler_init(batch, num_steps, sample_bb, sample_ang, sample_seq, noise=noise, uniforms=uniforms,
                                seed=seed, encoded=encoded, stream_to_host=True, graph=graph)
        for n in range(num_steps):
            smp.step(n)
        return smp.trajectory_to_host()

    # ------------------------------------------------------------------ training-style forward
    _IDEAL_BB = ((-0.525, 1.363, 0.0), (0.0, 0.0, 0.0), (1.526, 0.0, 0.0))  # N, CA, C in the backbone frame

    def _backbone_atoms(self, trans, rotmats):
        """N, CA, C from frames - the [:, :, :3] slice of data/all_atom.py:39-45 (to_atom37)."""
        return torch.einsum("blij,aj->blai", rotmats, self._ideal_bb.to(trans.dtype)) + trans[:, :, None, :]

    def forward(self, batch, *, noise=None):
        """Flow-matching losses (flow_model.py:111-227): returns the six-entry loss dict.
        With gradients disabled the denoiser runs the CUDA kernels; with gradients enabled it runs the
        autograd formulation (GAEncoder.forward_autograd).  `noise` optionally injects the corruption
        (keys t, trans_0, rotmats_0, angles_0, seqs_0_simplex, u_t, u_pred)."""
        num_batch, num_res = batch["aa"].shape
        dev = batch["aa"].device
        gen_b = batch["generate_mask"]
        gen_mask, res_mask = gen_b.long(), batch["res_mask"].long()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        rotmats_1, trans_1, angles_1, seqs_1, node_embed, edge_embed = self.encode(batch, autograd=need_grad)
        trans_1_c = trans_1
        seqs_1_simplex = self.seq_to_simplex(seqs_1)
        cfg = self._interpolant_cfg
        nz = noise or {}
        with torch.no_grad():
            t = nz["t"] if "t" in nz else torch.rand((num_batch, 1), device=dev)
            if "t" not in nz:
                t = t * (1 - 2 * cfg.min_t) + cfg.min_t
            if self.sample_structure:
                trans_0 = (nz["trans_0"] if "trans_0" in nz else torch.randn((num_batch, num_res, 3), device=dev)) * cfg.trans.sigma
                trans_0_c, _ = self.zero_center_part(trans_0, gen_mask, res_mask)
                trans_t = (1 - t[..., None]) * trans_0_c + t[..., None] * trans_1_c
                trans_t_c = torch.where(gen_b[..., None], trans_t, trans_1_c)
                rotmats_0 = nz["rotmats_0"] if "rotmats_0" in nz else uniform_so3(num_batch, num_res, device=dev)
                rotmats_t = so3_utils.geodesic_t(t[..., None], rotmats_1.contiguous(), rotmats_0.contiguous())
                rotmats_t = torch.where(gen_b[..., None, None], rotmats_t, rotmats_1)
                angles_0 = nz["angles_0"] if "angles_0" in nz else torus.tor_random_uniform(angles_1.shape, device=dev, dtype=angles_1.dtype)
                angles_t = torus.tor_geodesic_t(t[..., None], angles_1, angles_0)
                angles_t = torch.where(gen_b[..., None], angles_t, angles_1)
            else:
                trans_t_c, rotmats_t, angles_t = trans_1_c.detach().clone(), rotmats_1.detach().clone(), angles_1.detach().clone()
            if self.sample_sequence:
                seqs_0_simplex = nz["seqs_0_simplex"] if "seqs_0_simplex" in nz else self.k * torch.randn_like(seqs_1_simplex)
                seqs_t_simplex = ((1 - t[..., None]) * seqs_0_simplex) + (t[..., None] * seqs_1_simplex)
                seqs_t_simplex = torch.where(gen_b[..., None], seqs_t_simplex, seqs_1_simplex)
                p_t = F.softmax(seqs_t_simplex, dim=-1)
                seqs_t = categorical_from_uniform(p_t, nz["u_t"]) if "u_t" in nz else sample_from(p_t)
                seqs_t = torch.where(gen_b, seqs_t, seqs_1)
            else:
                seqs_t = seqs_1.detach().clone()

        if need_grad:
            pred_rotmats_1, pred_trans_1, pred_angles_1, pred_seqs_1_prob = self.ga_encoder.forward_autograd(
                t, rotmats_t, trans_t_c, angles_t, seqs_t, node_embed, edge_embed, gen_mask, res_mask)
        else:
            with torch.no_grad():
                pred_rotmats_1, pred_trans_1, pred_angles_1, pred_seqs_1_prob = self.ga_encoder(
                    t, rotmats_t, trans_t_c, angles_t, seqs_t, node_embed, edge_embed, gen_mask, res_mask)
        p_pred = F.softmax(pred_seqs_1_prob, dim=-1)
        pred_seqs_1 = categorical_from_uniform(p_pred, nz["u_pred"]) if "u_pred" in nz else sample_from(p_pred.detach())
        pred_seqs_1 = torch.where(gen_b, pred_seqs_1, torch.clamp(seqs_1, 0, 19))
        pred_trans_1_c = pred_trans_1

        norm_scale = 1 / (1 - torch.clamp(t[..., None], max=cfg.t_normalization_clip))
        gsum = torch.sum(gen_mask, dim=-1) + 1e-8
        trans_loss = torch.mean(torch.sum((pred_trans_1_c - trans_1_c) ** 2 * gen_mask[..., None], dim=(-1, -2)) / gsum)

        gt_rot_vf = so3_utils.calc_rot_vf(rotmats_t, rotmats_1)
        pred_rot_vf = _rot_vf_autograd(rotmats_t, pred_rotmats_1) if need_grad else so3_utils.calc_rot_vf(rotmats_t, pred_rotmats_1)
        rot_loss = torch.mean(torch.sum(((gt_rot_vf - pred_rot_vf) * norm_scale) ** 2 * gen_mask[..., None], dim=(-1, -2)) / gsum)

        gt_bb = self._backbone_atoms(trans_1_c, rotmats_1)
        pred_bb = self._backbone_atoms(pred_trans_1_c, pred_rotmats_1)
        bb_atom_loss = torch.mean(torch.sum((gt_bb - pred_bb) ** 2 * gen_mask[..., None, None], dim=(-1, -2, -3)) / gsum)

        seqs_loss = F.cross_entropy(pred_seqs_1_prob.reshape(-1, pred_seqs_1_prob.shape[-1]),
                                    torch.clamp(seqs_1, 0, 19).reshape(-1), reduction="none").view(pred_seqs_1_prob.shape[:-1])
        seqs_loss = torch.mean(torch.sum(seqs_loss * gen_mask, dim=-1) / gsum)

        aml = self._torsions_mask[pred_seqs_1.reshape(-1)].reshape(num_batch, num_res, -1)
        aml = torch.cat([aml, aml], dim=-1)
        aml = torch.logical_and(gen_b[..., None].bool(), aml)
        asum = torch.sum(aml, dim=(-1, -2)) + 1e-8
        gt_avf = torus.tor_logmap(angles_t, angles_1)
        pred_avf = torus.tor_logmap(angles_t, pred_angles_1)
        vec = lambda x: torch.cat([torch.sin(x), torch.cos(x)], dim=-1)
        angle_loss = torch.mean(torch.sum(((vec(gt_avf) - vec(pred_avf)) * norm_scale) ** 2 * aml, dim=(-1, -2)) / asum)
        torsion_loss = torch.mean(torch.sum((vec(pred_angles_1) - vec(angles_1)) ** 2 * aml, dim=(-1, -2)) / asum)
        return {"trans_loss": trans_loss, "rot_loss": rot_loss, "bb_atom_loss": bb_atom_loss, "seqs_loss": seqs_loss,
                "angle_loss": angle_loss, "torsion_loss": torsion_loss}


class EulerSampler:
    """Device-resident state of one FlowModel.sample call.  step(n) = one iteration of the reference loop
    (flow_model.py:287-343; the last one is :346-372) = ONE C-ABI call (pf_sampler_step: denoiser, post-processing into
    trajectory slot n, Euler update of the state) with no host synchronisation.  The iteration counter lives on the
    device and every argument is the same for all iterations, so the call is captured once in a CUDA graph and
    replayed (graph=True): 73 kernel launches per iteration become one graph launch."""

    HOST_CHUNK = 8   # trajectory slots per device->host transfer when streaming

    def __init__(self, model, batch, enc, noise, num_steps, flags, uniforms, seed, stream_to_host=False, graph=True):
        dev = batch["aa"].device
        B, L = batch["aa"].shape
        f32 = lambda x: x.to(torch.float32).contiguous()
        self.model, self.num_steps, self.flags, self.seed, self.dev = model, num_steps, flags, int(seed), dev
        self.B, self.L, self.k = B, L, model.k
        self.rot1, self.tr1, self.ang1 = f32(enc[0]), f32(enc[1]), f32(enc[2])
        self.seq1 = enc[3].to(torch.int64).contiguous()
        self.node_embed, self.edge_embed = f32(enc[4]), f32(enc[5])
        self.rot_t, self.tr_t, self.ang_t = (f32(noise["rotmats_0"]).clone(), f32(noise["trans_0"]).clone(),
                                             f32(noise["angles_0"]).clone())
        self.seq_t = noise["seqs_0"].to(torch.int64).contiguous().clone()
        self.sx_t = f32(noise["seqs_0_simplex"]).clone()
        self.tr0, self.sx0 = f32(noise["trans_0"]).clone(), f32(noise["seqs_0_simplex"]).clone()
        self.seq1_simplex = model.seq_to_simplex(self.seq1)
        self.gm_u8 = batch["generate_mask"].to(torch.uint8).contiguous()
        self.rm_f = f32(batch["res_mask"])
        self.tmask = torsions_mask.to(dev).contiguous()
        self.ts = torch.linspace(1.0e-2, 1.0, num_steps)                        # host fp32 (flow_model.py:280)
        self.ts_dev = self.ts.to(dev)
        self.t_dev = self.ts_dev[:, None].expand(num_steps, B).contiguous()      # row n: t of step n per complex
        if uniforms is not None:
            uniforms = f32(uniforms.to(dev))
            if uniforms.shape != (num_steps, 2, B, L):
                raise ValueError(f"uniforms must be [num_steps, 2, B, L], got {tuple(uniforms.shape)}")
        self.uniforms = uniforms
        self.traj = {"rotmats": torch.empty(num_steps, B, L, 3, 3, device=dev),
                     "trans": torch.empty(num_steps, B, L, 3, device=dev),
                     "angles": torch.empty(num_steps, B, L, 5, device=dev),
                     "seqs": torch.empty(num_steps, B, L, device=dev, dtype=torch.int64),
                     "seqs_simplex": torch.empty(num_steps, B, L, model.K, device=dev)}
        ga = model.ga_encoder
        self.weights, self._keep = ga.packed_weights()
        self.ws = ga.workspace(B, L, dev)
        self.pred = (torch.empty(B, L, 3, 3, device=dev), torch.empty(B, L, 3, device=dev),
                     torch.empty(B, L, 5, device=dev), torch.empty(B, L, 20, device=dev))
        self.gt = (self.rot1, self.tr1, self.ang1, self.seq1)
        # device-side bookkeeping of pf_sampler_step: [next iteration, scratch]; scratch t [B]
        self.step_dev = torch.zeros(2, dtype=torch.int32, device=dev)
        self.t_cur = torch.empty(B, device=dev)
        self._next = 0
        self.struct = self._make_struct()
        self.use_graph, self.graph, self.launches_per_step = bool(graph), None, None
        # Streaming of the clean trajectory to pinned host memory on a side stream while the loop runs (the
        # reference blocks on nine .cpu() calls per step, flow_model.py:313-314); torch's caching host allocator
        # recycles the pinned blocks of trajectories the caller has dropped.
        self.host = None
        if stream_to_host:
            self.host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in self.traj.items()}
            self.copy_stream = torch.cuda.Stream(device=dev)
            self._copied = 0

    def _make_struct(self):
        import ctypes
        s = _lib.Sampler()
        P = lambda t: t.data_ptr()
        s.weights = ctypes.addressof(self.weights)
        s.node_embed, s.edge_embed, s.res_mask = P(self.node_embed), P(self.edge_embed), P(self.rm_f)
        s.workspace, s.workspace_bytes = P(self.ws), self.ws.numel()
        s.rot1, s.trans1, s.ang1, s.seq1 = P(self.rot1), P(self.tr1), P(self.ang1), P(self.seq1)
        s.gen_mask, s.torsions_mask = P(self.gm_u8), P(self.tmask)
        s.trans0, s.simplex0 = P(self.tr0), P(self.sx0)
        s.rot_t, s.trans_t, s.ang_t, s.seq_t, s.simplex_t = (P(self.rot_t), P(self.tr_t), P(self.ang_t), P(self.seq_t),
                                                             P(self.sx_t))
        s.pred_rot, s.pred_trans, s.pred_ang, s.logits = (P(t) for t in self.pred)
        tj = self.traj
        s.traj_rot, s.traj_trans, s.traj_ang, s.traj_seq, s.traj_simplex = (
            P(tj["rotmats"]), P(tj["trans"]), P(tj["angles"]), P(tj["seqs"]), P(tj["seqs_simplex"]))
        s.ts = P(self.ts_dev)
        s.uniforms = P(self.uniforms) if self.uniforms is not None else None
        s.step, s.t_cur = P(self.step_dev), P(self.t_cur)
        s.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        s.num_steps, s.B, s.L = self.num_steps, self.B, self.L
        s.sample_bb, s.sample_ang, s.sample_seq = (int(bool(f)) for f in self.flags)
        s.simplex_k = float(self.k)
        return s

    def _flush_to_host(self, upto):
        """Enqueue the device->host copy of trajectory slots [self._copied, upto) behind the work issued so far."""
        if self.host is None or upto <= self._copied:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.copy_stream.wait_event(ev)
        with torch.cuda.stream(self.copy_stream):
            for k, v in self.traj.items():
                self.host[k][self._copied:upto].copy_(v[self._copied:upto], non_blocking=True)
        self._copied = upto

    def step(self, n):
        """Loop iteration n (time ts[n]); the clean prediction goes to trajectory slot n."""
        if not 0 <= n < self.num_steps:
            raise IndexError(f"iteration {n} outside [0, {self.num_steps})")
        with torch.cuda.device(self.dev):
            if n != self._next:                      # out-of-order call (tests, benchmarks): move the device counter
                self.step_dev[0:1].fill_(n)
            if self.use_graph:
                if self.graph is None:
                    # capture on a side stream with the raw API: torch.cuda.graph() would also synchronise the device,
                    # run the garbage collector and empty the caching allocator (0.5 s at the bench shape)
                    g = torch.cuda.CUDAGraph()
                    before = _lib.launch_count()
                    side, cur = torch.cuda.Stream(device=self.dev), torch.cuda.current_stream(self.dev)
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        g.capture_begin(capture_error_mode="thread_local")
                        try:
                            ops.sampler_step(self.struct, self.dev)
                        finally:
                            g.capture_end()
                    cur.wait_stream(side)
                    self.graph, self.launches_per_step = g, _lib.launch_count() - before
                self.graph.replay()
            else:
                ops.sampler_step(self.struct, self.dev)
        self._next = n + 1
        if self.host is not None and ((n + 1) % self.HOST_CHUNK == 0 or n == self.num_steps - 1):
            self._flush_to_host(n + 1)

    def step_unfused(self, n):
        """The same iteration as three C-ABI calls with host-side bookkeeping (pf_ga_encoder_forward, pf_denoise_post,
        pf_euler_step) - the cross-check of pf_sampler_step; not used by sample()."""
        sample_bb, sample_ang, sample_seq = self.flags
        u = self.uniforms
        ops.ga_encoder_forward(self.weights, self.t_dev[n], self.rot_t, self.tr_t, self.ang_t, self.seq_t,
                               self.node_embed, self.edge_embed, self.rm_f, self.ws, out=self.pred)
        tj = self.traj
        clean = (tj["rotmats"][n], tj["trans"][n], tj["angles"][n], tj["seqs"][n], tj["seqs_simplex"][n])
        ops.denoise_post(self.pred, self.gt, self.gm_u8, self.tmask, u[n, 0] if u is not None else None, self.seed,
                         2 * n, clean, self.k)
        if not sample_bb:
            clean[0].copy_(self.rot1); clean[1].copy_(self.tr1)
        if not sample_ang:
            clean[2].copy_(self.ang1)
        if not sample_seq:
            clean[3].copy_(self.seq1); clean[4].copy_(self.seq1_simplex)
        if n >= self.num_steps - 1:
            return
        d_t = float(self.ts[n + 1] - self.ts[n])
        ops.euler_step((self.rot_t, self.tr_t, self.ang_t, self.sx_t), clean[:4], (self.tr0, self.sx0), self.gt,
                       self.gm_u8, self.tmask, u[n, 1] if u is not None else None, self.seed, 2 * n + 1, d_t,
                       (self.rot_t, self.tr_t, self.ang_t, self.seq_t, self.sx_t), self.k)
        if not sample_bb:
            self.rot_t.copy_(self.rot1); self.tr_t.copy_(self.tr1)
        if not sample_ang:
            self.ang_t.copy_(self.ang1)
        if not sample_seq:
            self.seq_t.copy_(self.seq1)

    def trajectory_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.traj.values())

    def trajectory_to_host(self):
        """The reference's clean_traj: list of num_steps dicts of CPU tensors (flow_model.py:313-314,371-374),
        produced by ONE device->host copy per field after the loop instead of nine .cpu() calls per step."""
        if self.host is not None and self._copied == self.num_steps:
            self.copy_stream.synchronize()            # the streamed copies (issued during the loop) have landed
            host = self.host
        else:
            host = {k: v.cpu() for k, v in self.traj.items()}
        fixed = {"rotmats_1": self.rot1.cpu(), "trans_1": self.tr1.cpu(), "angles_1": self.ang1.cpu(),
                 "seqs_1": self.seq1.cpu()}
        out = []
        for n in range(self.num_steps):
            d = {k: host[k][n] for k in host}
            d.update(fixed)
            out.append(d)
        return out


def _rot_vf_autograd(mat_t, mat_1):
    """Differentiable Log(mat_t^T mat_1) for the training loss: the three-branch logarithm of data/so3_utils.py:167-254
    (theta ~ 0 Taylor, generic theta / (2 sin theta), theta ~ pi from (I + R) / 2 with the reference's sign choice) in
    torch ops, so the loss value equals the reference's.  The branch masks are piecewise constant; the square root of
    the theta ~ pi branch is clamped away from 0, where the reference's own backward produces the NaN gradients that
    train_ddp.py:139-142 zeroes."""
    rel = torch.einsum("...ji,...jk->...ik", mat_t, mat_1)
    v = torch.stack([rel[..., 2, 1] - rel[..., 1, 2], rel[..., 0, 2] - rel[..., 2, 0], rel[..., 1, 0] - rel[..., 0, 1]], dim=-1)
    sin_t = torch.linalg.norm(v, dim=-1) / 2.0
    cos_t = (rel[..., 0, 0] + rel[..., 1, 1] + rel[..., 2, 2] - 1.0) / 2.0
    th = torch.atan2(sin_t, cos_t)
    near_0 = torch.isclose(th, torch.zeros_like(th))
    near_pi = torch.isclose(th, torch.full_like(th, math.pi), atol=1e-2) & ~near_0
    generic = ~(near_0 | near_pi)
    one = torch.ones_like(th)
    pref = torch.where(generic, th / (2.0 * torch.where(generic, sin_t, one)), torch.zeros_like(th))
    pref = torch.where(near_0, 0.5 / (1.0 - th.detach() ** 2 / 6.0), pref)
    out = v * pref[..., None]
    if rel.is_cuda or bool(near_pi.any()):           # on the GPU always (no host synchronisation: graph-capturable)
        diag = torch.diagonal(rel, dim1=-2, dim2=-1)
        M = (torch.eye(3, device=rel.device, dtype=rel.dtype) + rel) / 2.0
        axis = torch.sqrt(torch.clamp((1.0 + diag) / 2.0, min=1e-12))
        row = torch.argmax(torch.linalg.norm(M.detach(), dim=-1), dim=-1)
        sgn = torch.sign(torch.take_along_dim(M.detach(), row[..., None, None], dim=-2).squeeze(-2))
        out = torch.where(near_pi[..., None], axis * th[..., None] * sgn, out)
    return out
