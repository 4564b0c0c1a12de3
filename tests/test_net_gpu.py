"""GPU: leaf-evaluation network (tcgen05 trunk, fused heads) against the fp32 oracle forward.

Tolerance (stated here as BASELINE.json asks): activations are bf16 between layers with fp32
accumulation.  On RANDOM-INIT weights - the worst case, logits have std ~11 (6x128) and tanh is
saturated - we require, over 256 random legal positions,
    policy KL(ref || ours): mean < 5e-3, max < 8e-2;   |dv|: mean < 3e-2, max < 0.3;
    argmax agreement >= 95 %.
A naive all-bf16 forward of the reference measures KL mean 1.2e-3 / max 2.0e-2, |dv| mean
1.2e-2 / max 0.20, so these bounds sit just above that noise floor."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import net as onet, rules as orules

pytestmark = pytest.mark.gpu


def positions(n, seed):
    rng = np.random.default_rng(seed)
    X = []
    for _ in range(n):
        p = orules.Position(0)
        for _ in range(int(rng.integers(0, 121))):
            e = np.flatnonzero(p.cells == 0)
            orules.play(p, int(e[int(rng.integers(0, len(e)))]))
        X.append(orules.encode(p))
    return np.stack(X).astype(np.float32)


def build(blocks, ch, seed=0, max_batch=512):
    import alphazero_gomoku_b200.network as mynet
    from alphazero_gomoku_b200.nn_engine import NetEngine
    torch.manual_seed(seed)
    net = mynet.AlphaZeroNet(n_res_blocks=blocks, channels=ch)
    # non-trivial BatchNorm statistics so that the folding is really tested
    g = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.8 + 0.4 * torch.rand(m.num_features, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    eng = NetEngine(blocks, ch, "cuda:0", max_batch=max_batch)
    eng.load_state_dict(sd)
    return sd, eng


@pytest.mark.parametrize("blocks,ch", [(1, 64), (1, 128), (1, 256)])
def test_trunk_layer_by_layer(blocks, ch):
    """Stem, then each 3x3 layer, against the fp32 oracle activations (bf16 rounding only)."""
    sd, eng = build(blocks, ch)
    X = positions(6, 3)
    with torch.no_grad():
        ref = [torch.relu(onet._bn(torch.nn.functional.conv2d(torch.from_numpy(X), sd["conv.weight"], padding=1), sd, "bn"))]
        h = ref[0]
        t = torch.relu(onet._bn(torch.nn.functional.conv2d(h, sd["res_blocks.0.conv1.weight"], padding=1), sd, "res_blocks.0.bn1"))
        ref.append(t)
        ref.append(torch.relu(onet._bn(torch.nn.functional.conv2d(t, sd["res_blocks.0.conv2.weight"], padding=1), sd, "res_blocks.0.bn2") + h))
    for n_layers, want in enumerate(ref):
        got = eng.trunk(torch.from_numpy(X).cuda(), n_layers).cpu()
        err = (got - want).abs().max().item()
        scale = want.abs().max().item()
        assert err <= 0.02 * scale + 0.02, (n_layers, err, scale)
    eng.close()


TOL = {(3, 64): (2e-4, 1.2e-3, 5e-3, 0.025, 0.97), (6, 128): (1e-3, 1.5e-2, 1.5e-2, 0.2, 0.97)}


@pytest.mark.parametrize("blocks,ch", [(3, 64), (6, 128)])
def test_policy_value_tolerance(blocks, ch):
    import alphazero_gomoku_b200.network as mynet
    from alphazero_gomoku_b200.nn_engine import NetEngine
    torch.manual_seed(0)
    net = mynet.AlphaZeroNet(n_res_blocks=blocks, channels=ch)      # the reference's random init (seed 0)
    sd = net.state_dict()
    eng = NetEngine(blocks, ch, "cuda:0", max_batch=128)              # 256 positions -> two chunks
    eng.load_state_dict(sd)
    X = positions(256, 5)
    with torch.no_grad():
        lo, v_ref = onet.forward(sd, torch.from_numpy(X))
        p_ref = torch.softmax(lo, dim=1).numpy()
    probs, values, logits = eng.forward(torch.from_numpy(X).cuda(), want_logits=True)
    probs, values = probs.cpu().numpy(), values.cpu().numpy()
    assert np.allclose(probs.sum(1), 1.0, atol=1e-4)
    kl = onet.policy_kl(p_ref, probs)
    dv = np.abs(values - v_ref.numpy())
    agree = float((probs.argmax(1) == p_ref.argmax(1)).mean())
    print(f"{blocks}x{ch}: KL mean {kl.mean():.2e} max {kl.max():.2e}  |dv| mean {dv.mean():.2e} max {dv.max():.2e}  argmax {agree:.3f}")
    # stated bf16 tolerance = about twice what this kernel measures (round 1: 6x128 KL mean 4.3e-4 / max 6.7e-3,
    # |dv| mean 6.8e-3 / max 0.10; 3x64 KL mean 9.4e-5 / max 5.3e-4, |dv| max 0.011)
    kl_mean, kl_max, dv_mean, dv_max, min_agree = TOL[(blocks, ch)]
    assert kl.mean() < kl_mean and kl.max() < kl_max
    assert dv.mean() < dv_mean and dv.max() < dv_max
    assert agree >= min_agree
    eng.close()


def test_policy_value_tolerance_10x256():
    """BASELINE config 5 network (10 blocks x 256 channels, streaming-weights trunk kernel).  Random-init
    logits have std ~109 here, so the softmax is one-hot and a near-tie can flip the argmax: the
    bound is on agreement and on the value head (|dv| max 5.7e-2 for naive bf16)."""
    import alphazero_gomoku_b200.network as mynet
    from alphazero_gomoku_b200.nn_engine import NetEngine
    torch.manual_seed(0)
    net = mynet.AlphaZeroNet(n_res_blocks=10, channels=256)
    sd = net.state_dict()
    eng = NetEngine(10, 256, "cuda:0", max_batch=64)
    eng.load_state_dict(sd)
    X = positions(64, 9)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        lo, v_ref = onet.forward(sd, torch.from_numpy(X))
        p_ref = torch.softmax(lo, dim=1).numpy()
    probs, values, logits = eng.forward(torch.from_numpy(X).cuda(), want_logits=True)
    probs, values, logits = probs.cpu().numpy(), values.cpu().numpy(), logits.cpu().numpy()
    rel = np.abs(logits - lo.numpy()).max() / np.abs(lo.numpy()).max()
    agree = float((probs.argmax(1) == p_ref.argmax(1)).mean())
    dv = np.abs(values - v_ref.numpy())
    print(f"10x256: max logit err / max |logit| {rel:.3e}  argmax {agree:.3f}  |dv| mean {dv.mean():.2e} max {dv.max():.2e}")
    assert rel < 0.03 and agree >= 0.9 and dv.max() < 0.3
    eng.close()


def test_golden_reference_outputs():
    """Same seed-0 weights as the reference: outputs vs the fixture the reference produced."""
    import alphazero_gomoku_b200.network as mynet
    z = load_golden("net_outputs.npz")
    torch.manual_seed(0)
    model = mynet.PyTorchModel(board_size=15, n_res_blocks=6, channels=128, device="cuda:0")
    probs, values = model.predict(z["X"])
    assert probs.shape == (24, 225) and values.shape == (24, 1) and probs.dtype == np.float32
    kl = onet.policy_kl(z["6x128/probs"], probs)
    kl_mean, kl_max, dv_mean, dv_max, _ = TOL[(6, 128)]
    dv = np.abs(values - z["6x128/values"])
    print(f"golden 6x128: KL mean {kl.mean():.2e} max {kl.max():.2e}  |dv| mean {dv.mean():.2e} max {dv.max():.2e}")
    assert kl.mean() < kl_mean and kl.max() < kl_max
    assert dv.mean() < dv_mean and dv.max() < dv_max


def test_search_with_real_network_on_device():
    """MCTS drop-in with this package's PyTorchModel (device path) against the oracle search fed
    by the SAME network outputs: identical priors in, identical visit counts out."""
    import alphazero_gomoku_b200 as m
    import alphazero_gomoku_b200.network as mynet
    from oracle.search import Search

    class Gomoku:
        pass

    torch.manual_seed(0)
    model = mynet.PyTorchModel(board_size=15, n_res_blocks=3, channels=64, device="cuda:0")

    class Recorder:            # oracle evaluator: the CUDA network through its numpy API
        def predict(self, X):
            return model.predict(X)

    mcts = m.MCTS(Gomoku, 200, model, cpuct=1.0, batch_size=32, add_dirichlet_noise=False)
    orc = Search(0, 200, Recorder(), cpuct=1.0, queue_len=32, noise=False)
    pos = orules.Position(0)

    class G:
        pass
    for move in range(3):
        g = G()
        g.board = pos.cells.reshape(15, 15).copy(); g.current_player = pos.player
        g.last_move = None if pos.last < 0 else divmod(pos.last, 15); g.move_history = [None] * pos.plies
        pi = mcts.run(g, pos.plies)
        want = orc.run(pos, pos.plies)
        assert np.array_equal(mcts.last_visits, orc.Nv[pos.key()].astype(np.int32)), move
        assert np.array_equal(pi, want)
        orules.play(pos, int(np.argmax(pi)))


def test_outputs_do_not_depend_on_batch_shape():
    """A row of the heads GEMM depends on that board's features only and the trunk handles boards independently:
    a position evaluates to identical bits whatever batch it travels in (the search's sharding and pipelining
    invariance rests on this)."""
    sd, eng = build(2, 128, seed=3, max_batch=1024)
    X = torch.from_numpy(positions(24, seed=11)).cuda()
    small_p, small_v = eng.forward(X)                        # 24 boards: one partly filled 128-board tile
    big = X.repeat(30, 1, 1, 1)                              # 720 boards: six tiles on six SMs
    big_p, big_v = eng.forward(big)
    assert torch.equal(small_p, big_p[:24]) and torch.equal(small_p, big_p[-24:])
    assert torch.equal(small_v, big_v[:24]) and torch.equal(small_v, big_v[-24:])
    one_p, one_v = eng.forward(X[5:6])
    assert torch.equal(one_p[0], small_p[5]) and torch.equal(one_v[0], small_v[5])
    eng.close()


@pytest.mark.parametrize("ch", [64, 256])
def test_network_without_residual_blocks(ch):
    """n_res_blocks = 0: stem -> unfused 1x1 head convolutions (head1_kernel) -> heads GEMM; the only path
    on which 64/128-channel networks use the unfused head kernel."""
    sd, eng = build(0, ch, seed=4, max_batch=64)
    X = positions(48, 21)
    with torch.no_grad():
        lo, v_ref = onet.forward(sd, torch.from_numpy(X))
        p_ref = torch.softmax(lo, dim=1).numpy()
    probs, values = eng.forward(torch.from_numpy(X).cuda())
    kl = onet.policy_kl(p_ref, probs.cpu().numpy())
    dv = np.abs(values.cpu().numpy() - v_ref.numpy())
    assert kl.max() < 2e-2 and dv.max() < 0.1, (kl.max(), dv.max())
    eng.close()
