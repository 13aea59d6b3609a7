"""CPU: host-side logic - batch schema, collate, embedders against the reference's golden outputs,
categorical sampler against the oracle, checkpoint key handling, fresh-init conventions."""
import numpy as np
import pytest
import torch

from oracle import pepflow_oracle as orc
from tests.conftest import load_golden, rel_err


def test_synthetic_batch_schema():
    from pepflowww_b200.pep_dataloader import PaddingCollate, SyntheticPepDataset, synthetic_batch
    ds = SyntheticPepDataset(num_complexes=3, len_pocket=10, len_peptide=4, seed=1)
    item = ds[0]
    assert item["aa"].shape == (14,) and item["pos_heavyatom"].shape == (14, 15, 3)
    assert item["generate_mask"].sum() == 4 and not item["generate_mask"][:10].any()
    assert item["chain_nb"][:10].eq(1).all() and item["chain_nb"][10:].eq(0).all()
    ca = item["pos_heavyatom"][10:, 1]
    assert ca.mean(0).abs().max() < 1e-4          # peptide CA centroid at the origin
    b = PaddingCollate(eight=True)([ds[0], ds[1]])
    assert b["aa"].shape == (2, 16) and b["res_mask"].sum() == 28 and (b["aa"][:, 14:] == 21).all()
    b2 = synthetic_batch(2, 18, 5, seed=7)
    g = load_golden("encode")
    assert torch.equal(b2["aa"], g["aa"]) and torch.equal(b2["pos_heavyatom"], g["pos_heavyatom"])


def test_embedders_match_reference_golden(state_dict):
    from pepflowww_b200.config import load_config
    from pepflowww_b200.flow_model import FlowModel
    cfg, _ = load_config()
    model = FlowModel(cfg.model).eval()
    model.load_state_dict(state_dict)
    g = load_golden("encode")
    batch = {k: g[k] for k in ("aa", "pos_heavyatom", "mask_heavyatom", "res_nb", "chain_nb", "generate_mask",
                               "res_mask", "torsion_angle", "torsion_angle_mask")}
    with torch.no_grad():
        r1, x1, a1, s1, node, edge = model.encode(batch, autograd=True)   # the torch formulations (training path) on CPU
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.encode(batch)                        # the kernel path refuses CPU tensors instead of falling back
        model.edge_embedder.chunk_bytes = 1 << 16      # force the row-chunked path too
        edge_chunked = model.edge_embedder.forward_autograd(batch["aa"], batch["res_nb"], batch["chain_nb"], batch["pos_heavyatom"],
                                           batch["mask_heavyatom"], structure_mask=~batch["generate_mask"],
                                           sequence_mask=~batch["generate_mask"])
    assert rel_err(r1, g["rotmats_1"]) < 1e-6
    assert rel_err(node, g["node_embed"]) < 2e-5
    assert rel_err(edge, g["edge_embed"]) < 2e-5
    assert rel_err(edge_chunked, g["edge_embed"]) < 2e-5


def test_categorical_matches_oracle():
    from pepflowww_b200.layers import categorical_from_uniform
    rng = np.random.default_rng(0)
    p = torch.softmax(torch.from_numpy(rng.standard_normal((4, 50, 20)).astype(np.float32)) * 3, -1)
    u = torch.from_numpy(rng.random((4, 50), dtype=np.float32))
    assert torch.equal(categorical_from_uniform(p, u), orc.categorical_from_uniform(p, u))


def test_process_dic_and_load(state_dict):
    from pepflowww_b200.config import load_config
    from pepflowww_b200.flow_model import FlowModel
    from pepflowww_b200.utils import process_dic
    cfg, _ = load_config()
    ddp_style = {"module." + k: v for k, v in state_dict.items()}
    model = FlowModel(cfg.model)
    missing = model.load_state_dict(process_dic(ddp_style))
    assert not missing.missing_keys and not missing.unexpected_keys


def test_fresh_init_final_layers_are_zero():
    from pepflowww_b200.config import load_config
    from pepflowww_b200.flow_model import FlowModel
    cfg, _ = load_config()
    m = FlowModel(cfg.model)
    t = m.ga_encoder.trunk
    for name in ("ipa_0.linear_out", "post_tfmr_0", "node_transition_0.linear_3", "bb_update_0.linear",
                 "edge_transition_0.final_layer"):
        mod = t
        for part in name.split("."):
            mod = mod[part] if isinstance(mod, torch.nn.ModuleDict) else getattr(mod, part)
        assert mod.weight.abs().sum() == 0 and mod.bias.abs().sum() == 0
    assert abs(float(t["ipa_0"].head_weights[0]) - 0.541324854612918) < 1e-6


def test_time_frequencies_bitwise():
    from pepflowww_b200.utils_time import get_time_embedding
    t = torch.tensor([0.01, 0.5, 1.0])
    assert torch.equal(get_time_embedding(t, 128, 2056), orc.time_embedding(t))


def test_embedder_kernel_constants_match_the_concatenated_first_layers():
    """The fused embedder kernels take the embedding lookups already pushed through the first Linear of their MLPs
    (EdgeEmbedder._kernel_constants / NodeEmbedder._kernel_constants).  On CPU: the table form equals the reference's
    concatenate-then-Linear form (models_con/edge.py:104-105, models_con/node.py:98-99)."""
    import torch

    from pepflowww_b200.edge import EdgeEmbedder
    from pepflowww_b200.node import NodeEmbedder
    torch.manual_seed(0)
    ee = EdgeEmbedder(64, 15)
    torch.nn.init.normal_(ee.aapair_to_distcoef.weight, std=0.02)
    c = ee._kernel_constants()
    assert [tuple(t.shape) for t in c] == [(484, 225), (484, 64), (65, 64), (225, 64), (64,), (64, 64), (64,), (64, 64),
                                           (26, 64), (64,), (64, 64), (64,), (64, 64), (64,)]
    n = 50
    aa_pair, rel = torch.randint(0, 484, (n,)), torch.randint(0, 65, (n,))
    same = torch.randint(0, 2, (n, 1)).float()
    f_d, f_h = torch.rand(n, 64), torch.randn(n, 26)
    ref = ee.out_mlp[0](torch.cat([ee.aa_pair_embed(aa_pair), ee.relpos_embed(rel) * same, f_d, f_h], dim=-1))
    got = c[1][aa_pair] + same * c[2][rel] + f_d @ c[7] + f_h @ c[8] + c[9]
    assert torch.allclose(got, ref, atol=1e-5)
    assert torch.allclose(c[0], torch.nn.functional.softplus(ee.aapair_to_distcoef.weight))

    ne = NodeEmbedder(128, 15)
    c = ne._kernel_constants()
    assert [tuple(t.shape) for t in c] == [(22, 256), (990, 256), (39, 256), (256, 128), (128,), (128, 128), (128,),
                                           (128, 128), (128,)]
    aa = torch.randint(0, 22, (n,))
    crd, dih = torch.randn(n, 45), torch.randn(n, 39)
    place = torch.zeros(n, 22, 45)
    place[torch.arange(n), aa] = crd
    ref = ne.mlp[0](torch.cat([ne.aatype_embed(aa), place.reshape(n, 990), dih], dim=-1))
    w1c = c[1].reshape(22, 45, 256)
    got = c[0][aa] + torch.einsum("nk,nkc->nc", crd, w1c[aa]) + dih @ c[2]
    assert torch.allclose(got, ref, atol=1e-5)


def test_heavyatom_mask_and_pdb_writer_round_trip():
    """get_heavyatom_mask equals the reference's table gather (golden), and save_pdb's fixed-column records parse back
    to the atoms that went in (names per residue type, chain split, serial numbers, coordinates to 1e-3)."""
    from pepflowww_b200 import torsion, writers
    from pepflowww_b200.constants import AA, heavyatom_names
    from tests.conftest import load_golden
    g = load_golden("reconstruction")
    mask = torsion.get_heavyatom_mask(g["aa"])
    assert mask.dtype == torch.bool and (mask == g["mask15"]).all()
    i = 0
    L = g["aa"].shape[1]
    data = dict(chain_nb=g["chain_nb"][i], aa=g["aa"][i], resseq=g["res_nb"][i], icode=[" "] * L,
                chain_id=["A" if c == 0 else "B" for c in g["chain_nb"][i].tolist()],
                pos_heavyatom=torch.nn.functional.pad(g["pos14"][i], (0, 0, 0, 1)), mask_heavyatom=mask[i])
    text = writers.save_pdb(data)
    lines = text.splitlines()
    assert all(len(ln) == 80 for ln in lines[:-1]) and lines[-1] == "END   "
    assert sum(ln.startswith("TER") for ln in lines) == 2
    atoms = writers.parse_pdb_atoms(text)
    assert len(atoms) == int(mask[i].sum())
    k = 0
    for r in range(L):
        for a, name in enumerate(heavyatom_names(AA(int(g["aa"][i, r])))):
            if name == "" or not mask[i, r, a]:
                continue
            serial, nm, resname, chain, resseq, xyz = atoms[k]
            assert nm == name and resname == AA(int(g["aa"][i, r])).name and resseq == int(g["res_nb"][i, r])
            assert chain == ("A" if g["chain_nb"][i, r] == 0 else "B")
            assert max(abs(x - float(y)) for x, y in zip(xyz, g["pos14"][i, r, a])) < 6e-4
            k += 1
    assert [a[0] for a in atoms[:5]] == [1, 2, 3, 4, 5]


def test_autograd_rot_vf_matches_reference_on_every_branch():
    """The differentiable SO(3) logarithm of the training loss equals the reference's calc_rot_vf on the manifold fixture
    (theta = 0, tiny, generic, within 1e-2 of pi, exactly pi) and its backward stays finite on all of them."""
    from pepflowww_b200.flow_model import _rot_vf_autograd
    from tests.conftest import load_golden, rel_err
    g = load_golden("manifold")
    target = g["target"].clone().requires_grad_(True)
    vf = _rot_vf_autograd(g["base"], target)
    assert rel_err(vf.detach(), g["rot_vf"]) < 1e-6
    vf.square().sum().backward()
    assert torch.isfinite(target.grad).all() and float(target.grad.abs().sum()) > 0


def test_autograd_denoiser_matches_reference_and_reaches_every_parameter(state_dict):
    """GAEncoder.forward_autograd (the gradient path of train_ddp.py) against the reference's GAEncoder output (golden,
    padded and unpadded batch) on CPU tensors; every ga_encoder parameter receives a finite gradient."""
    from pepflowww_b200.config import load_config
    from pepflowww_b200.flow_model import FlowModel
    from tests.conftest import circ_err, load_golden, rel_err
    cfg, _ = load_config()
    m = FlowModel(cfg.model)
    m.load_state_dict(state_dict)
    keys = ("t", "rotmats_t", "trans_t", "angles_t", "seqs_t", "node_embed", "edge_embed", "generate_mask", "res_mask")
    for tag in ("ga_encoder_a", "ga_encoder_b"):
        g = load_golden(tag)
        R, x, ang, logits = m.ga_encoder.forward_autograd(*[g[k] for k in keys])
        assert rel_err(R.detach(), g["out_rotmats"]) < 2e-5 and rel_err(x.detach(), g["out_trans"]) < 2e-5
        assert circ_err(ang.detach(), g["out_angles"]) < 2e-5 and rel_err(logits.detach(), g["out_logits"]) < 2e-5
    (R.sum() + x.sum() + torch.sin(ang).sum() + logits.square().sum()).backward()
    grads = [p.grad for p in m.ga_encoder.parameters()]
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads)
    with torch.enable_grad(), pytest.raises(RuntimeError):     # the kernel path refuses to run under autograd
        m.ga_encoder(*[g[k] for k in keys])


def test_save_samples_pipeline_on_host(tmp_path, monkeypatch):
    """sample.save_samples_sc / _bb (models_con/sample.py:68-120): file set, chain split, context residues untouched,
    generated residues replaced - with the device geometry substituted by the oracle so the host logic runs on CPU."""
    from pepflowww_b200 import constants, sample, writers
    from pepflowww_b200.pep_dataloader import synthetic_batch
    batch = synthetic_batch(3, 10, 4, seed=2)
    T = constants.rigid_tables("cpu")
    B, L = batch["aa"].shape
    g = torch.Generator().manual_seed(1)
    q = torch.nn.functional.normalize(torch.randn(B, L, 4, generator=g), dim=-1)
    samples = {"rotmats": orc.quat_to_rot(q), "trans": torch.randn(B, L, 3, generator=g) * 5,
               "angles": torch.rand(B, L, 5, generator=g) * 6.28, "seqs": torch.randint(0, 20, (B, L), generator=g),
               "batch": batch}
    samples["seqs"] = torch.where(batch["generate_mask"], samples["seqs"], batch["aa"])

    def side(s):
        pos14, _, _ = orc.full_atom_reconstruction(T, s["rotmats"], s["trans"], s["angles"], s["seqs"])
        return torch.nn.functional.pad(pos14, (0, 0, 0, 1)), constants.restype_to_heavyatom_masks[s["seqs"]]

    monkeypatch.setattr(sample, "_device_of", lambda s: torch.device("cpu"))
    monkeypatch.setattr(sample.torsion, "reconstruct_side_chains", side)
    monkeypatch.setattr(sample.geometry, "reconstruct_backbone",
                        lambda R, t, aa, c, r, m: orc.reconstruct_backbone(T, R, t, aa, c, r, m))
    for fn, n_gen_atoms in ((sample.save_samples_sc, None), (sample.save_samples_bb, 4)):
        paths = fn(samples, str(tmp_path / fn.__name__))
        assert [p.split("/")[-1] for p in paths] == ["sample_0.pdb", "sample_1.pdb", "sample_2.pdb", "gt.pdb"]
        atoms = writers.parse_pdb_atoms(open(paths[0]).read())      # gt.pdb is complex 0 (replicas in the reference)
        gt = writers.parse_pdb_atoms(open(paths[-1]).read())
        assert {a[3] for a in atoms} == {"A", "B"}
        ctx = lambda rows: [a[1:] for a in rows if a[3] == "B"]
        assert ctx(atoms) == ctx(gt)                                  # pocket (context) residues: written unchanged
        pep = [a for a in atoms if a[3] == "A"]
        assert len({a[4] for a in pep}) == 4
        if n_gen_atoms:
            assert len(pep) == 4 * n_gen_atoms and {a[1] for a in pep} == {"N", "CA", "C", "O"}
        else:
            assert len(pep) == int(constants.restype_to_heavyatom_masks[samples["seqs"][0][batch["generate_mask"][0]]].sum())


def test_inference_metrics_arithmetic():
    """sample_metrics (models_con/inference.py:76-78): RMSDs and recovery are taken over generated residues only."""
    from pepflowww_b200.inference import sample_metrics
    B, L = 2, 6
    gm = torch.zeros(B, L, dtype=torch.bool)
    gm[:, 4:] = True
    eye = torch.eye(3).expand(B, L, 3, 3)
    final = {"trans": torch.zeros(B, L, 3), "trans_1": torch.zeros(B, L, 3), "rotmats": eye.clone(), "rotmats_1": eye.clone(),
             "seqs": torch.zeros(B, L, dtype=torch.long), "seqs_1": torch.zeros(B, L, dtype=torch.long)}
    final["trans"][:, :4] += 100.0          # context residues never count
    final["trans"][:, 4:, 0] = 3.0
    final["seqs"][0, 5] = 7
    m = sample_metrics(final, {"generate_mask": gm})
    assert abs(m["tran"] - 3.0) < 1e-5 and m["rot"] == 0.0 and abs(m["aar"] - 0.75) < 1e-6 and m["len"] == 4


def test_padding_collate_matches_reference():
    """PaddingCollate (pepflow/utils/data.py:19-78) against the reference's output on ragged items with list / string
    fields: same keys (common keys only), same padding values (aa -> 21, chain_id / icode -> ' ', others 0), same
    res_mask, same transposed list-of-tuples layout for the string lists, with and without the multiple-of-8 rounding."""
    import json
    from tests.golden.make_golden_collate import items
    from pepflowww_b200.pep_dataloader import PaddingCollate
    g = load_golden("collate")
    for tag, eight in (("e8", True), ("e1", False)):
        b = PaddingCollate(eight=eight)(items())
        want = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + "_")}
        assert {k.replace("__json", "") for k in want} == set(b), (sorted(want), sorted(b))
        for k, v in want.items():
            if k.endswith("__json"):
                ref = json.loads(bytes(v.numpy().tolist()).decode())
                ours = json.loads(json.dumps(b[k[:-6]]))
                assert ours == ref, k
            else:
                assert b[k].dtype == v.dtype and torch.equal(b[k], v), k
    assert PaddingCollate(no_padding={"resseq"}).no_padding == {"resseq"}


def test_benchdata_generators_equal_the_products(state_dict):
    """bench.py's reference arm builds its workload from benchdata/ (no product import): same complexes, same weights."""
    import benchdata
    from pepflowww_b200.constants import restype_to_heavyatom_masks, torsions_mask
    from pepflowww_b200.pep_dataloader import synthetic_batch
    a = benchdata.synthetic_batch(3, 20, 5, seed=2, first_index=4)
    b = synthetic_batch(3, 20, 5, seed=2, first_index=4)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    sd = benchdata.reference_state_dict(114514)
    assert set(sd) == set(state_dict)
    for k in sd:
        assert torch.equal(sd[k], state_dict[k]), k
    assert torch.equal(benchdata.torsions_mask, torsions_mask)
    from benchdata.synthetic import restype_to_heavyatom_masks as hm
    assert torch.equal(hm, restype_to_heavyatom_masks)


def test_checkpoint_round_trip_and_reference_style_checkpoint(tmp_path, state_dict):
    """train -> save -> load must work on torch >= 2.6 (weights_only default) and a checkpoint written by the
    reference's DDP loop ('module.' keys, config pickled as easydict.EasyDict, iteration = last completed) must load."""
    import pickle
    import sys
    import types

    from pepflowww_b200.config import load_config
    from pepflowww_b200.flow_model import FlowModel
    from pepflowww_b200.train import checkpoint_dict, get_optimizer, get_scheduler
    from pepflowww_b200.utils import load_checkpoint, process_dic
    cfg, _ = load_config()
    net = FlowModel(cfg.model)
    net.load_state_dict(state_dict)
    opt = get_optimizer(cfg.train.optimizer, net)
    sch = get_scheduler(cfg.train.scheduler, opt)
    path = str(tmp_path / "ckpt.pt")
    torch.save(checkpoint_dict(cfg, net, opt, sch, 41), path)
    ck = load_checkpoint(path)
    assert ck["iteration"] == 41 and type(ck["config"]) is dict and ck["config"]["train"]["batch_size"] == cfg.train.batch_size
    net2 = FlowModel(cfg.model)
    net2.load_state_dict(process_dic(ck["model"]))
    for k, v in net2.state_dict().items():
        assert torch.equal(v, state_dict[k]), k
    # the same file also loads with plain torch.load defaults minus the config (tensors + containers only)
    assert torch.load(path, weights_only=True)["iteration"] == 41

    # reference-style: EasyDict config from a module this image does not have, DDP-prefixed keys
    fake = types.ModuleType("easydict")

    class EasyDict(dict):
        pass

    EasyDict.__module__, EasyDict.__qualname__ = "easydict", "EasyDict"
    fake.EasyDict = EasyDict
    sys.modules["easydict"] = fake
    try:
        ref_path = str(tmp_path / "model1.pt")
        torch.save({"config": EasyDict(model=EasyDict(encoder=1)), "model": {"module." + k: v for k, v in state_dict.items()},
                    "optimizer": {}, "scheduler": {}, "iteration": 7}, ref_path, pickle_protocol=pickle.DEFAULT_PROTOCOL)
    finally:
        del sys.modules["easydict"]
    assert "easydict" not in sys.modules
    ck = load_checkpoint(ref_path)
    assert ck["config"].model.encoder == 1 and ck["iteration"] == 7
    net3 = FlowModel(cfg.model)
    net3.load_state_dict(process_dic(ck["model"]))
    assert "easydict" not in sys.modules


def test_reference_arm_does_not_import_the_product(tmp_path):
    """`bench.py --impl reference` must not import pepflowww_b200 nor map libpepflow_b200.so."""
    import subprocess
    import sys

    from tests.conftest import ROOT
    code = (
        "import runpy, sys\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-batch', '1', '--pocket', '12', '--peptide', '4']\n"
        "try:\n"
        "    runpy.run_path('bench.py', run_name='__main__')\n"
        "except SystemExit:\n"
        "    pass\n"
        "bad = [m for m in sys.modules if m.startswith('pepflowww_b200')]\n"
        "assert not bad, bad\n"
        "assert 'libpepflow' not in open('/proc/self/maps').read()\n"
        "print('REFERENCE_ARM_CLEAN')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "REFERENCE_ARM_CLEAN" in r.stdout
    import json
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["gpu_launches"] == 0


def test_dihedral_rounding_model():
    """The embedder kernels compute dihedral_from_four_points (geometry.py:296-313) with explicitly rounded operations in
    the order PyTorch's CPU kernels use (pf_geom.cuh::dihedral4): cross = fma(a1, b2, -(a2 b1)), norm = sqrt(fma chain),
    sums left to right.  This test pins that model against the installed torch build, bit for bit, on the layouts the
    reference calls it with - if it ever fails, the kernel and the CPU reference differ again by the acos / sign
    conditioning (1e-3 on ~1000 pairs per complex)."""
    f32 = np.float32

    def fma(x, y, z):
        return (x.astype(np.float64) * y.astype(np.float64) + z.astype(np.float64)).astype(f32)

    def cross(a, b):
        c = lambda x, y, z, w: fma(x, y, -(z * w).astype(f32))
        return np.stack([c(a[..., 1], b[..., 2], a[..., 2], b[..., 1]), c(a[..., 2], b[..., 0], a[..., 0], b[..., 2]),
                         c(a[..., 0], b[..., 1], a[..., 1], b[..., 0])], -1)

    def norm(u):
        return np.sqrt(fma(u[..., 2], u[..., 2], fma(u[..., 1], u[..., 1], (u[..., 0] * u[..., 0]).astype(f32))))

    def dot3(a, b):
        m = (a * b).astype(f32)
        return ((m[..., 0] + m[..., 1]).astype(f32) + m[..., 2]).astype(f32)

    def model(p0, p1, p2, p3):
        v0, v1, v2 = (p2 - p1).astype(f32), (p0 - p1).astype(f32), (p3 - p2).astype(f32)
        u1, u2 = cross(v0, v1), cross(v0, v2)
        n1, n2 = (u1 / norm(u1)[..., None]).astype(f32), (u2 / norm(u2)[..., None]).astype(f32)
        return np.clip(dot3(n1, n2), f32(-0.999999), f32(0.999999)), np.sign(dot3(cross(v1, v2), v0))

    def torch_ops(p0, p1, p2, p3):
        v0, v1, v2 = p2 - p1, p0 - p1, p3 - p2
        u1 = torch.cross(v0, v1, dim=-1)
        n1 = u1 / torch.linalg.norm(u1, dim=-1, keepdim=True)
        u2 = torch.cross(v0, v2, dim=-1)
        n2 = u2 / torch.linalg.norm(u2, dim=-1, keepdim=True)
        return ((n1 * n2).sum(-1).clamp(min=-0.999999, max=0.999999), torch.sign((torch.cross(v1, v2, dim=-1) * v0).sum(-1)))

    g = torch.Generator().manual_seed(1)
    for N, L in ((2, 140), (3, 37), (1, 8)):
        pos = torch.randn(N, L, 3, 3, generator=g) * 6
        pN, pCA, pC = pos[:, :, 0], pos[:, :, 1], pos[:, :, 2]
        rows = lambda x: x[:, :, None].expand(N, L, L, 3)
        cols = lambda x: x[:, None, :].expand(N, L, L, 3)
        for args in ((rows(pC), cols(pN), cols(pCA), cols(pC)), (rows(pN), rows(pCA), rows(pC), cols(pN)),
                     (pCA[:, :-1], pC[:, :-1], pN[:, 1:], pCA[:, 1:])):
            x, s = torch_ops(*args)
            xm, sm = model(*[a.numpy() for a in args])
            assert np.array_equal(xm, x.numpy()) and np.array_equal(sm, s.numpy())


def test_nccl_overlap_summary_interval_arithmetic():
    """train.nccl_overlap_summary (the source of the NCCL kernel time / overlap figures of DESIGN.md section 7): NCCL kernel
    intervals intersected with the union of compute-kernel intervals, per iteration."""
    from types import SimpleNamespace as NS
    from pepflowww_b200.train import nccl_overlap_summary

    def ev(name, a, b, dev="DeviceType.CUDA"):
        return NS(name=name, device_type=dev, time_range=NS(start=a, end=b))

    events = [ev("gemm_a", 0, 100), ev("gemm_b", 50, 150),                  # union [0, 150]
              ev("gemm_c", 300, 400),
              ev("ncclDevKernel_AllReduce_Sum_f32_RING_LL", 120, 320),      # 200 us: overlaps [120,150] + [300,320] = 50 us
              ev("ncclDevKernel_AllReduce_Sum_f32_RING_LL", 1000, 1100),    # 100 us, no compute underneath
              ev("Memcpy DtoD", 1000, 1100),                                # copies do not count as compute
              ev("cpu_op", 0, 5000, dev="DeviceType.CPU")]
    out = nccl_overlap_summary(NS(events=lambda: events), iters=2)
    assert abs(out["nccl_kernel_ms_per_iter"] - 0.150) < 1e-9               # (200 + 100) us / 2 iterations
    assert out["nccl_launches_per_iter"] == 1.0
    assert abs(out["overlapped_with_compute"] - 50.0 / 300.0) < 1e-9
    assert out["compute_kernel_launches_per_iter"] == 1.5
    assert nccl_overlap_summary(NS(events=lambda: events[:3]), iters=1)["overlapped_with_compute"] is None
