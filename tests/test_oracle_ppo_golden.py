"""oracle/ppo.py and learner.MpiAdam against known answers of the reference's OWN code (tests/golden/ref_ppo_kat.npz,
made by scripts/make_ppo_golden.py from R/agents/ppo_imitation.py:68-93 and R/stable_baselines/common/mpi_adam.py:8-70)."""
import io
import json
import zipfile

import numpy as np
import torch

from oracle import ppo as oppo
from tests import common

KAT = np.load(common.GOLD + "/ref_ppo_kat.npz")


def test_oracle_gae_is_the_reference_function():
    for c, (T, N) in enumerate(KAT["gae/cases"]):
        g = lambda k: KAT["gae/%d/%s" % (c, k)]
        adv, ret = oppo.add_vtarg_and_adv(g("rewards"), g("vpred"), g("nextvpreds"), g("episode_starts"), int(N), 0.95, 0.95)
        assert np.array_equal(adv, g("adv")) and np.array_equal(ret, g("tdlamret"))      # same float32 arithmetic, bit for bit
        assert adv.dtype == np.float32


def test_oracle_adam_is_the_reference_update():
    for tag in ("adam1", "adam2"):
        theta = KAT[tag + "/theta0"].astype(np.float32)
        m, v, step = np.zeros_like(theta), np.zeros_like(theta), 0
        for i, (g, lr) in enumerate(zip(KAT[tag + "/grads"], KAT[tag + "/lrs"])):
            if tag == "adam2":
                g = (g + KAT[tag + "/other_grads"][i]) / 2          # Allreduce(SUM) / comm.Get_size()
            theta, m, v, step = oppo.mpi_adam_update(theta, m, v, step, g, lr, epsilon=1e-5)
            np.testing.assert_allclose(theta, KAT[tag + "/thetas"][i], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(m, KAT[tag + "/exp_avg"], rtol=2e-6, atol=1e-8)
        np.testing.assert_allclose(v, KAT[tag + "/exp_avg_sq"], rtol=2e-6, atol=1e-9)


def test_learner_mpi_adam_against_the_reference_vectors():
    from openroborl_b200.learner import MpiAdam
    tag = "adam1"                                                  # single rank (the 2-rank path runs over gloo in test_learner_cpu.py)
    sizes = KAT[tag + "/shapes"]
    theta0 = KAT[tag + "/theta0"]
    params, off = [], 0
    for n in sizes:
        params.append(torch.tensor(theta0[off:off + n].copy()))
        off += n
    opt = MpiAdam(params, epsilon=1e-5)
    opt.sync()
    for i, (g, lr) in enumerate(zip(KAT[tag + "/grads"], KAT[tag + "/lrs"])):
        opt.update(torch.tensor(g), float(lr))
        np.testing.assert_allclose(opt.getflat().numpy(), KAT[tag + "/thetas"][i], rtol=3e-6, atol=1e-7)
    np.testing.assert_allclose(opt.exp_avg.numpy(), KAT[tag + "/exp_avg"], rtol=3e-6, atol=1e-8)


def test_saved_zip_has_the_layout_the_reference_loader_reads(tmp_path):
    """BaseRLModel._load_from_file / load_parameters (R/stable_baselines/common/base_class.py:552-690): a zip with `data`
    (json), `parameter_list` (json list of variable names, graph order) and `parameters` (npz keyed by those names); the
    reference rebuilds an OrderedDict by walking parameter_list and feeds it to the policy's variables by name."""
    from openroborl_b200.learner import PPOLearner, PARAM_ORDER
    from openroborl_b200.rollout import MlpGaussianPolicy, load_policy_zip
    z = np.load(common.GOLD + "/policy_laikago_pace.npz")
    params = {k.rsplit("_", 1)[0] + "/" + k.rsplit("_", 1)[1]: z[k] for k in z.files if k.rsplit("_", 1)[-1] in ("w", "b")}
    pol = MlpGaussianPolicy(params, "cpu")
    path = str(tmp_path / "model.zip")
    PPOLearner(pol).save(path, data={"n_envs": 1})
    with zipfile.ZipFile(path) as f:
        assert set(f.namelist()) == {"data", "parameters", "parameter_list"}
        data = json.loads(f.read("data"))
        names = json.loads(f.read("parameter_list"))
        npz = np.load(io.BytesIO(f.read("parameters")))
        # the walk of base_class.py:669-675
        loaded = {n: npz[n] for n in names}
    assert data["gamma"] == 0.95 and data["lam"] == 0.95 and data["clip_param"] == 0.2
    want = ["model/%s:0" % k for k in PARAM_ORDER] + ["model/q/w:0", "model/q/b:0"]
    assert names == want                                           # the variable order of the reference's own policy zips
    for k in PARAM_ORDER:
        assert loaded["model/%s:0" % k].dtype == np.float32 and np.array_equal(loaded["model/%s:0" % k], params[k])
    assert loaded["model/q/w:0"].shape == (256, 12)
    back = load_policy_zip(path)                                   # and this repo's reader returns the same arrays
    assert all(np.array_equal(back[k], params[k]) for k in PARAM_ORDER)
