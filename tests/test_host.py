"""CPU tests of the host-side mirror: configuration, masks, schedule modules, library surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLDEN, histogram_file, make_model
from oracle import hd_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports exactly what include/hierdiff_b200.h declares."""
    from hierdiff_b200 import build, native
    build.build()
    header = open(os.path.join(ROOT, "include", "hierdiff_b200.h")).read()
    declared = set(re.findall(r"HD_API\s+[\w\s\*]+?\b(hd_\w+)\s*\(", header))
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    L = ctypes.CDLL(native.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert native.lib().hd_abi_version() == native.ABI_VERSION


def test_layout_sizes_match_reference_parameter_count():
    from hierdiff_b200 import native
    for n_layers, expect in [(4, 3959574), (6, 5935386)]:    # SURVEY.md 8b [probe]: whole model
        cfg = native.HdConfig(n_layers, 2, 256, 9, 1, 1, 30.0, 0.0, 10.0, 0)
        n = native.lib().hd_weight_count(cfg)
        assert n == O.lib().hdo_weight_count(ctypes.byref(O.make_config(n_layers)))
        assert n + 3077 == expect         # + the gamma network's 3077 parameters
        assert native.lib().hd_packed_bytes(cfg) > 4 * n
    bad = native.HdConfig(4, 2, 128, 9, 1, 1, 30.0, 0.0, 10.0, 0)
    assert native.lib().hd_weight_count(bad) < 0 and "hidden_nf" in native.last_error()


def test_state_dict_keys_and_shapes_equal_reference(tmp_path):
    model = make_model(tmp_path, n_layers=2)
    sd = model.state_dict()
    cfg = O.make_config(2)
    ref = dict(O.egnn_shapes(cfg))
    ref.update({"buffer": (1,), "gamma.gamma_0": (1,), "gamma.gamma_1": (1,), "gamma.l1.weight": (1, 1),
                "gamma.l1.bias": (1,), "gamma.l2.weight": (1024, 1), "gamma.l2.bias": (1024,),
                "gamma.l3.weight": (1, 1024), "gamma.l3.bias": (1,)})
    assert {k: tuple(v.shape) for k, v in sd.items()} == ref
    # flat parameter order of the EGNN == order of the native flat buffer
    names = ["dynamics.egnn." + n for n, _ in model.dynamics.egnn.named_parameters()]
    assert names == O.egnn_key_order(cfg)


def test_gamma_network_matches_reference_fixture(tmp_path):
    g = np.load(os.path.join(GOLDEN, "gamma.npz"))
    model = make_model(tmp_path, n_layers=1)
    with torch.no_grad():
        got = model.gamma(torch.from_numpy(g["t"]).view(-1, 1))[:, 0].numpy()
    assert np.abs(got - g["gamma_batched"]).max() < 1e-5
    model2 = make_model(tmp_path, n_layers=1, noise_schedule="polynomial_2")
    assert np.array_equal(model2.gamma.gamma.numpy(), g["gamma_poly2_table"])
    with torch.no_grad():
        got2 = model2.gamma(torch.from_numpy(g["t"]).view(-1, 1))[:, 0].numpy()
    assert np.array_equal(got2, g["gamma_poly2"])


def test_nodes_dist_reproduces_reference_draws(tmp_path):
    g = np.load(os.path.join(GOLDEN, "nodes_dist.npz"))
    model = make_model(tmp_path, n_layers=1)
    torch.manual_seed(0)
    assert model.nodes_dist.sample(64) == g["seed0_64"].tolist()
    assert model.nodes_dist.sample(7) == g["then_7"].tolist()


def test_masks_roundtrip_and_rejection():
    from hierdiff_b200.utils import check_edge_index, check_edge_mask, masks_from_sizes, sizes_from_node_mask
    sizes = [3, 5, 1]
    nm, em = masks_from_sizes(sizes, 5, "cpu")
    assert nm.shape == (3, 5, 1) and em.shape == (3, 5, 5)
    assert sizes_from_node_mask(nm, 3, 5).tolist() == sizes
    check_edge_mask(em, torch.tensor(sizes), 3, 5)
    assert int(em[1].sum()) == 20 and not em[0, 3:].any() and not em[0, :, 3:].any()
    bad = nm.clone()
    bad[0, 0] = False
    with pytest.raises(NotImplementedError):
        sizes_from_node_mask(bad, 3, 5)
    em2 = em.clone()
    em2[1, 0, 0] = True
    with pytest.raises(NotImplementedError):
        check_edge_mask(em2, torch.tensor(sizes), 3, 5)
    from hierdiff_b200 import EGNN_dynamics_QM9
    dyn = EGNN_dynamics_QM9(in_node_nf=9, context_node_nf=0, n_dims=3, hidden_nf=256)
    rows, cols = dyn.get_adj_matrix(4, 2)
    assert rows[:5].tolist() == [0, 0, 0, 0, 1] and cols[:5].tolist() == [0, 1, 2, 3, 0] and rows[16] == 4
    check_edge_index((rows, cols), 2, 4)
    with pytest.raises(NotImplementedError):
        check_edge_index((cols, rows), 2, 4)


def test_config_composition(tmp_path):
    from hierdiff_b200.config import instantiate, load_config
    conf = tmp_path / "conf"
    (conf / "model").mkdir(parents=True)
    (conf / "sample").mkdir()
    hist = histogram_file(tmp_path)
    (conf / "sample.yaml").write_text(
        "defaults:\n  - model: tiny\n  - sample: default\n  - override hydra/job_logging: colorlog\n"
        "checkpoint: /nowhere/diffusion.ckpt\nhydra:\n  run:\n    dir: x/${now:%Y}\n")
    (conf / "sample" / "default.yaml").write_text("batch_size: 2\nnum_batches: 16\n")
    from hierdiff_b200.config import default_model_cfg
    import yaml
    node = {"_target_": "train_module.diffusion_qm9.DiffusionQM9",
            "cfg": dict(default_model_cfg(n_layers=1, analyze=hist))}
    node["cfg"]["dynamics"] = dict(node["cfg"]["dynamics"])
    node["cfg"]["pre_noise"] = dict(node["cfg"]["pre_noise"])
    (conf / "model" / "tiny.yaml").write_text(yaml.safe_dump(node))
    cfg = load_config(str(conf), "sample", ["sample.batch_size=5", "model.cfg.timesteps=50"])
    assert cfg.sample.batch_size == 5 and cfg.sample.num_batches == 16
    assert cfg.checkpoint == "/nowhere/diffusion.ckpt" and "hydra" not in cfg
    model = instantiate(cfg.model, cfg=cfg, _recursive_=False)
    assert model.T == 50 and model.dynamics.egnn.n_layers == 1
    assert model.cfg.sample.batch_size == 5            # root keys are visible on the model cfg, as with hydra
    assert model.cfg.dynamics.in_node_nf == 9          # mutated like the reference (diffusion_qm9.py:47,90)
    with pytest.raises(NotImplementedError):
        instantiate({"_target_": "somewhere.Else"}, cfg=cfg)


def test_product_path_fails_loudly_without_cuda(tmp_path):
    """No CPU fallback: sampling on a CPU device raises instead of silently running PyTorch."""
    from hierdiff_b200 import native
    model = make_model(tmp_path, n_layers=1, timesteps=5)
    with pytest.raises(native.NativeError):
        model.sample(2, torch.device("cpu"))
    nm = torch.ones(1, 4, 1, dtype=torch.bool)
    em = ~torch.eye(4, dtype=torch.bool).unsqueeze(0)
    with pytest.raises(native.NativeError):
        model.phi(torch.zeros(1, 4, 11), torch.zeros(1, 1), nm, em, None)


def _cpu_guard():
    raise ValueError("no CUDA device in this test environment")


def test_unsupported_options_raise(tmp_path):
    from hierdiff_b200 import EGNN, EGNN_dynamics_QM9
    with pytest.raises(NotImplementedError):
        EGNN(9, 1, 256, sin_embedding=True)
    with pytest.raises(NotImplementedError):
        EGNN_dynamics_QM9(9, 0, 3, mode="gnn_dynamics")
    dyn = EGNN_dynamics_QM9(9, 2, 3, hidden_nf=256)     # context conditioning is built: 9 + 2 input channels
    assert dyn.egnn.embedding.weight.shape == (256, 11) and dyn.egnn.embedding_out.weight.shape == (11, 256)
    with pytest.raises(ValueError):                     # ... and refuses a call without its context
        dyn.forward_sizes(torch.zeros(1), torch.zeros(1, 2, 11).cuda() if torch.cuda.is_available()
                          else _cpu_guard(), torch.ones(1, dtype=torch.int32))


def test_pocket_model_keeps_reference_state_dict(tmp_path):
    """cfg.pocket adds pocket_embed (diffusion_qm9.py:55-56) and nothing else."""
    plain = make_model(tmp_path, n_layers=1, timesteps=5)
    pocket = make_model(tmp_path, n_layers=1, timesteps=5, pocket=True)
    extra = set(pocket.state_dict()) - set(plain.state_dict())
    assert extra == {"pocket_embed.weight"} and tuple(pocket.state_dict()["pocket_embed.weight"].shape) == (21, 8)
    with pytest.raises(ValueError):
        plain._check_pocket_cond([torch.zeros(1, 2, dtype=torch.long)] * 4, 1)


REF_CONF = "/root/reference/endiffusion/conf"


@pytest.mark.skipif(not os.path.isdir(REF_CONF), reason="the reference checkout only exists in the build container")
def test_reference_config_tree_loads_unchanged():
    """conf/sample.yaml of the reference (defaults list -> model/ddpmgblur.yaml, sample/default.yaml, analyze/GEOM.yaml)
    composes unchanged and its `_target_: train_module.diffusion_qm9.DiffusionQM9` resolves to the native mirror."""
    from hierdiff_b200 import DiffusionQM9
    from hierdiff_b200.config import instantiate, load_config
    cfg = load_config(REF_CONF, "sample", ["sample.batch_size=4"])
    assert cfg.model["_target_"] == "train_module.diffusion_qm9.DiffusionQM9"
    assert cfg.sample.batch_size == 4 and cfg.sample.num_batches == 16
    model = instantiate(cfg.model, cfg=cfg, _recursive_=False)
    assert isinstance(model, DiffusionQM9) and model.T == 1000
    egnn = model.dynamics.egnn
    assert (egnn.n_layers, egnn.hidden_nf, egnn.inv_sublayers) == (6, 256, 2)      # ddpmgblur.yaml:21-37
    assert sum(p.numel() for p in model.dynamics.parameters()) == 5935386 - sum(p.numel() for p in model.gamma.parameters())
    assert len(model.nodes_dist.n_nodes) == 67                                     # conf/analyze/GEOM.yaml


def test_header_is_plain_c_and_links(tmp_path):
    """include/hierdiff_b200.h compiles as C99 and examples/c_host.c links against the library and runs (host-only
    entry points: sizes and capability queries, no GPU needed)."""
    import subprocess
    from hierdiff_b200 import native
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_host")
    libdir = os.path.dirname(native.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "c_host.c"), "-L", libdir, "-lhierdiff_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.check_output([exe], text=True)
    assert "parameters 3956497 floats" in out and "strict=1" in out and "hidden_nf=128 ->" in out


def test_merged_sample_batches_bookkeeping(tmp_path):
    """merge_batches: the pooled chains see the sizes / contexts of the reference's per-batch draws
    (diffusion_qm9.py:349,:431-432), sorted by size and cut at max_chain_molecules, and every molecule comes back in
    its batch-order slot.  The chain itself is replaced by a recorder, so this runs without a GPU."""
    model = make_model(tmp_path, 1, timesteps=4, context_node_nf=1)
    model.merge_batches, model.max_chain_molecules = True, 5
    chains = []

    def fake_chain(sample_n, device, z_T=None, context=None):
        chains.append(list(sample_n))
        B, N = len(sample_n), max(sample_n)
        assert context.shape == (B, N, 1)
        x = torch.tensor(sample_n, dtype=torch.float32).view(B, 1, 1).expand(B, N, 3).clone()
        h = context.expand(B, N, 1).repeat(1, 1, 8).clone()
        return x, h

    model.sample_padded = fake_chain
    torch.manual_seed(11)
    out, names = model.sample_batches(4, 3, "cuda", context_range=[0.5, -2.0])
    torch.manual_seed(11)
    want = [int(n) for _ in range(3) for n in model.nodes_dist.sample(4)]   # the reference's draws, batch by batch
    assert names == [] and [r["x"].shape[0] for r in out] == want
    assert [len(c) for c in chains] == [4, 4, 4] and sum(chains, []) == sorted(want)
    for k, r in enumerate(out):
        c = [0.5, -2.0][(k // 4) % 2]
        assert torch.all(r["x"] == want[k]) and torch.all(r["h"] == c) and torch.all(r["context"] == c)
        assert r["h"].shape == (want[k], 8) and r["context"].shape == (want[k], 1)


def test_argument_checks_of_the_ragged_entry_point_need_no_gpu():
    """hd_dynamics_forward_ragged validates its shape arguments before touching the device: a bound outside
    [0, B*N], more than 4096 molecules on a tensor-core engine or N > 128 return HD_E_INVALID with a message."""
    from hierdiff_b200 import native
    L = native.lib()
    hc = native.HdConfig(1, 2, 256, 9, 1, 1, 30.0, 0.0, 10.0, 0)   # EGNN(n_layers=1, ...) as egnn.hd_config() builds it
    one = ctypes.c_void_p(16)   # never dereferenced: every call below fails validation first

    def call(B, N, live, engine=native.ENGINE_TC_STRICT):
        return L.hd_dynamics_forward_ragged(ctypes.byref(hc), one, one, one, None, 0, one, B, N, live, one, one, None,
                                            engine, None)

    assert call(2, 3, 7) == -1 and b"live_rows" in L.hd_last_error()
    assert call(2, 3, -1) == -1
    assert call(2, 129, 0) == -1 and b"N" in L.hd_last_error()


def test_stage2_and_loss_entry_points_validate_before_touching_the_device():
    """hd_egcl_* / hd_linear_forward / hd_loss_*: sizes come from the configuration alone, bad arguments return
    HD_E_INVALID / HD_E_UNSUPPORTED with a message - none of this needs a GPU."""
    from hierdiff_b200 import native
    L = native.lib()
    full = native.HdEgclConfig(256, 256, 1, 1, 30.0, 1)     # gcl_full_*: hidden edge features, attention, edge update
    plain = native.HdEgclConfig(256, 1, 0, 1, 30.0, 0)      # gcl_edge / gcl_denoise
    H = 256
    n_full = (H * (2 * H + 1 + H) + H) + (H * H + H) + (H * (H + 1 + H) + H) + (H * H + H) + (H * 2 * H + H) + (H * H + H) \
        + (H * H + H) + H + (H + 1)
    assert L.hd_egcl_weight_count(ctypes.byref(full)) == n_full
    assert L.hd_egcl_packed_bytes(ctypes.byref(full)) > 0 and L.hd_egcl_packed_bytes(ctypes.byref(plain)) == 0
    assert L.hd_egcl_workspace_bytes(ctypes.byref(full), 10, 100) > 100 * H * 4 * 4
    one = ctypes.c_void_p(16)   # never dereferenced

    def egcl(cfg, row, col, bits, sizes, B, N, n_nodes, n_edges, edge_attr=one, packed=None, engine=native.ENGINE_FP32):
        return L.hd_egcl_forward(ctypes.byref(cfg), one, packed, one, one, edge_attr, row, col, bits, None, None, sizes, B, N,
                                 n_nodes, n_edges, one, one, one, one, engine, None)

    assert egcl(full, None, one, 32, one, 2, 5, 0, 0) == -1 and b"dense" in L.hd_last_error()       # dense list with a col
    assert egcl(full, one, None, 32, None, 0, 0, 10, 4) == -1 and b"explicit" in L.hd_last_error()   # list without col
    assert egcl(full, one, one, 16, None, 0, 0, 10, 4) == -1 and b"index_bits" in L.hd_last_error()
    assert egcl(full, one, one, 64, None, 0, 0, 10, 4, edge_attr=None) == -1                          # hidden features need edge_attr
    assert egcl(plain, None, None, 32, one, 2, 5, 0, 0, engine=native.ENGINE_TC_STRICT) == -3        # no tensor-core path: De = 1
    assert b"tensor-core" in L.hd_last_error()
    assert L.hd_linear_forward(one, 4, 0, one, None, 8, 0, one, None) == -1
    lc = native.HdLossConfig(1000, 1, 0, 5, 3, 1.0, 1.0, 0.0)
    assert L.hd_loss_terms(ctypes.byref(lc), *([one] * 3), one, None, None, None, *([one] * 6), 2, 5, 8, one, one, one, None,
                           None) == -1                                                            # t0_always without z_0
    lc.int_nf = 9                                                                                  # more columns than F
    assert L.hd_loss_terms(ctypes.byref(lc), *([one] * 13), 2, 5, 8, one, one, one, None, None) == -1
    assert L.hd_loss_prepare(one, one, one, 0, 5, 8, 1.0, 1.0, 0.0, 1, one, None, None) == -1


def test_ragged_rows_rule():
    """SamplingLoop.ragged_rows_pay: the hint is set only for batches beyond one wave of node-GEMM CTAs with at least
    a quarter of padding."""
    from hierdiff_b200.sampling import SamplingLoop
    pay = SamplingLoop.ragged_rows_pay
    assert not pay([40] * 64, 64, 40)                  # headline shape: no padding
    assert not pay([3] * 63 + [40], 64, 40)            # much padding but 2560 rows: one wave either way
    assert pay([5] * 255 + [40], 256, 40)
    assert not pay([35] * 256, 256, 40)                # 12 % padding only


def test_load_checkpoint_reads_a_lightning_style_file(tmp_path):
    """sampler.py:26-34: ``torch.load(ckpt)['state_dict']`` with the ``model.`` prefix stripped.  The reference's file is
    a lightning checkpoint whose ``hyper_parameters`` entry is an OmegaConf object (diffusion_qm9.py:40): neither
    ``weights_only=True`` nor a plain unpickle (omegaconf is not installed) can read it; tensors must still load."""
    import sys
    import types
    from hierdiff_b200.sampler import load_checkpoint, read_state_dict
    model = make_model(tmp_path, 1)
    other = make_model(tmp_path, 1, seed=5)
    fake = types.ModuleType("omegaconf_like")

    DictConfig = type("DictConfig", (dict,), {"__module__": "omegaconf_like", "__qualname__": "DictConfig"})
    fake.DictConfig = DictConfig
    sys.modules["omegaconf_like"] = fake
    try:
        ckpt = {"state_dict": {"model." + k: v for k, v in other.state_dict().items()},
                "hyper_parameters": DictConfig(cfg=DictConfig(timesteps=1000)), "epoch": 7,
                "pytorch-lightning_version": "1.4.9"}
        path = str(tmp_path / "diffusion.ckpt")
        torch.save(ckpt, path)
    finally:
        del sys.modules["omegaconf_like"]           # the class is NOT importable when the file is read
    state = read_state_dict(path)
    assert set(state) == set(other.state_dict())
    load_checkpoint(model, path)
    for k, v in other.state_dict().items():
        assert torch.equal(model.state_dict()[k], v), k
    plain = str(tmp_path / "plain.ckpt")             # a tensors-only file takes the weights_only=True path
    torch.save({"state_dict": other.state_dict()}, plain)
    assert set(read_state_dict(plain)) == set(other.state_dict())


def test_masks_are_validated_on_every_call():
    """utils.sizes_from_masks keeps no pointer-keyed cache: a second mask pair of the same shape (possibly at a
    recycled address) is validated and reduced again."""
    from hierdiff_b200.utils import masks_from_sizes, sizes_from_masks
    nm, em = masks_from_sizes([3, 5], 5, "cpu")
    B, N, s1 = sizes_from_masks(nm.view(10, 1), em.view(50, 1), None, 10)
    assert (B, N, s1.tolist()) == (2, 5, [3, 5])
    nm.copy_(masks_from_sizes([4, 1], 5, "cpu")[0])      # same storage, same shape, other content
    em.copy_(masks_from_sizes([4, 1], 5, "cpu")[1])
    assert sizes_from_masks(nm.view(10, 1), em.view(50, 1), None, 10)[2].tolist() == [4, 1]


def test_staged_reference_copy_is_intact():
    """oracle/_ref (oracle/stage_ref.py) is what the GPU parity tests and bench.py's reference legs run: when present
    it must verify against its manifest and reproduce a committed fixture."""
    from oracle import ref_runner, stage_ref
    if not stage_ref.available():
        pytest.skip("oracle/_ref not staged in this checkout (python oracle/stage_ref.py)")
    assert stage_ref.verify()
    g = np.load(os.path.join(GOLDEN, "sample_poly_l1.npz"))
    ref = ref_runner.make_reference(int(g["n_layers"]), int(g["T"]), noise_schedule="polynomial_2")
    x, h = ref_runner.sample_padded(ref, g["sizes"], "cpu", int(g["sample_seed"]))
    assert np.array_equal(x, g["x"]) and np.array_equal(h, g["h"])
