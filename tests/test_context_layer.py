"""Context plugin surface: ported semantics of the reference's test/test_context_space.py,
test_context_sampler.py, test_context_selector.py, test_search_space_encoding.py plus the golden
sampler stream extracted from the reference notebooks (tests/golden/notebook_goldens.json)."""
import json
import os

import numpy as np
import pytest

from carl_b200 import spaces
from carl_b200.context import (
    CategoricalContextFeature, ContextSampler, ContextSpace, CustomSelector, NormalFloatContextFeature,
    RandomSelector, RoundRobinSelector, StaticSelector, UniformFloatContextFeature, UniformIntegerContextFeature,
)
from carl_b200.context.search_space_encoding import FeatureSpace, search_space_to_config_space
from carl_b200.context.utils import get_context_bounds

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_goldens.json")))

context_space_dict = {
    "gravity": UniformFloatContextFeature("gravity", lower=0.1, upper=np.inf, default_value=9.8),
    "masscart": UniformFloatContextFeature("masscart", lower=0.1, upper=10, default_value=1.0),
    "masspole": UniformFloatContextFeature("masspole", lower=0.01, upper=1, default_value=0.1),
    "length": UniformFloatContextFeature("length", lower=0.05, upper=5, default_value=0.5),
    "force_mag": UniformFloatContextFeature("force_mag", lower=1, upper=100, default_value=10.0),
    "tau": UniformFloatContextFeature("tau", lower=0.002, upper=0.2, default_value=0.02),
}
DEFAULT = {"gravity": 9.8, "masscart": 1, "masspole": 0.1, "length": 0.5, "force_mag": 10, "tau": 0.02}


def test_insert_defaults_and_default_context():
    cs = ContextSpace(context_space_dict)
    assert cs.insert_defaults({}) == DEFAULT
    assert cs.get_default_context() == DEFAULT
    assert cs.insert_defaults({"gravity": 3.0})["gravity"] == 3.0


def test_bounds_and_space_types():
    cs = ContextSpace(context_space_dict)
    assert cs.get_lower_and_upper_bound("length") == (0.05, 5)
    assert type(cs.to_gymnasium_space(as_dict=False)) is spaces.Box
    assert type(cs.to_gymnasium_space(as_dict=True)) is spaces.Dict
    other = ContextSpace({
        "gravity": UniformFloatContextFeature("gravity", lower=0.1, upper=np.inf, default_value=9.8),
        "masscart": UniformIntegerContextFeature("masscart", lower=1, upper=10, default_value=1),
    })
    other.to_gymnasium_space()


def test_verify_context():
    cs = ContextSpace(context_space_dict)
    assert cs.verify_context({"hihi": 39, "gravity": 3}) is False
    assert cs.verify_context({"masscart": -10}) is False
    assert cs.verify_context({"masscart": 2.0}) is True


def test_sample_contexts():
    cs = ContextSpace(context_space_dict)
    c = cs.sample_contexts(["gravity"], size=1)
    assert cs.verify_context(c)
    many = cs.sample_contexts(["gravity"], size=10)
    assert len(many) == 10 and all(cs.verify_context(c) for c in many)
    many = cs.sample_contexts(None, size=10)
    assert len(many) == 10 and all(cs.verify_context(c) for c in many)
    with pytest.raises(ValueError):
        cs.sample_contexts(["false_feature"], size=0)


def _sampler():
    cspace = ContextSpace({"gravity": UniformFloatContextFeature("gravity", lower=1, upper=10, default_value=9.8)})
    dist = {"gravity": NormalFloatContextFeature("gravity", mu=9.8, sigma=0.0, default_value=9.8, upper=20, lower=1)}
    return cspace, dist


def test_sampler_init_and_sigma0():
    cspace, dist = _sampler()
    ContextSampler(context_distributions=dist, context_space=cspace, seed=0, name="TestSampler")
    ContextSampler(context_distributions=list(dist.values()), context_space=cspace, seed=0, name="TestSampler")
    with pytest.raises(ValueError):
        ContextSampler(context_distributions=0, context_space=cspace, seed=0, name="TestSampler")
    s = ContextSampler(context_distributions=dist, context_space=cspace, seed=0)
    c = s.sample_contexts(n_contexts=3)
    assert len(c) == 3 and c[0]["gravity"] == 9.8
    c = s.sample_contexts(n_contexts=1)
    assert len(c) == 1 and c[0]["gravity"] == 9.8


def _ant_space():
    from carl_b200.envs.brax import CARLBraxAnt

    return CARLBraxAnt.get_context_space()


def test_sampler_golden_gravity_stream():
    """examples/sample_contexts_with_brax.ipynb cell 5 (seed 0, Normal gravity)."""
    s = ContextSampler([NormalFloatContextFeature("gravity", mu=9.8, sigma=1, upper=50, lower=0)],
                       context_space=_ant_space(), seed=0)
    got = s.sample_contexts(n_contexts=5)
    want = {int(k): v for k, v in GOLD["ant_gravity_normal_seed0_n5"].items()}
    assert got == want


def test_sampler_golden_categorical_drawn_first():
    """examples/brax_with_goals.ipynb cell 1: target_direction (categorical) precedes target_distance."""
    from carl_b200.envs.brax import directions

    s = ContextSampler(
        [NormalFloatContextFeature("target_distance", mu=9.8, sigma=1, upper=50, lower=0),
         CategoricalContextFeature("target_direction", choices=directions)],
        context_space=_ant_space(), seed=0)
    got = s.sample_contexts(n_contexts=5)
    want = {int(k): v for k, v in GOLD["ant_target_seed0_n5"].items()}
    assert got == want


def test_sampler_golden_two_normals():
    """examples/brax_with_goals.ipynb cell 4: x column then y column from one RandomState."""
    cs = ContextSpace({
        "goal_position_x": UniformFloatContextFeature("goal_position_x", -np.inf, np.inf, default_value=0.45),
        "goal_position_y": UniformFloatContextFeature("goal_position_y", -np.inf, np.inf, default_value=-0.05),
    })
    s = ContextSampler(
        [NormalFloatContextFeature("goal_position_x", mu=9.8, sigma=1, upper=50, lower=0),
         NormalFloatContextFeature("goal_position_y", mu=9.8, sigma=1, upper=50, lower=0)], context_space=cs, seed=0)
    got = s.sample_contexts(5)
    for k, v in GOLD["pusher_goal_xy_seed0_n5"].items():
        assert got[int(k)]["goal_position_x"] == v["goal_position_x"]
        assert got[int(k)]["goal_position_y"] == v["goal_position_y"]


def test_sample_context_table_matches_dicts():
    feats = [UniformFloatContextFeature("gravity", 5, 15), UniformFloatContextFeature("length", 0.25, 1.0)]
    cs = ContextSpace(context_space_dict)
    a = ContextSampler(feats, cs, seed=3).sample_contexts(7)
    names = list(DEFAULT)
    t = ContextSampler(feats, cs, seed=3).sample_context_table(7, names)
    for i in range(7):
        assert [a[i][n] for n in names] == list(t[i])


def _contexts():
    c = {"dt": 0.03, "gravity": 10.0, "m": 1.0, "l": 1.8}
    return {k: dict(c, m=float(i)) for i, k in enumerate("abc")}


def test_round_robin_and_batch_equivalence():
    s = RoundRobinSelector(_contexts())
    assert s.context_id is None and s.context_key is None
    ids = []
    for _ in range(7):
        s.select()
        ids.append(s.context_id)
    assert ids == [0, 1, 2, 0, 1, 2, 0] and s.n_calls == 7
    s2 = RoundRobinSelector(_contexts())
    assert list(s2.select_batch(4)) + list(s2.select_batch(3)) == ids and s2.n_calls == 7
    assert s2.context_id == ids[-1]


def test_static_random_custom_selectors():
    s = StaticSelector(_contexts())
    assert list(s.select_batch(3)) == [0, 0, 0]
    assert s.select() == _contexts()["a"]
    np.random.seed(0)
    r = RandomSelector(_contexts())
    a = []
    for _ in range(5):
        r.select()
        a.append(int(r.context_id))
    np.random.seed(0)
    r2 = RandomSelector(_contexts())
    assert list(r2.select_batch(5)) == a and r2.n_calls == 5
    assert set(a) <= {0, 1, 2}

    def fn(inst):
        cid = 1 if inst.n_calls == 0 else 0
        return inst.contexts[inst.contexts_keys[cid]], cid

    c = CustomSelector(_contexts(), fn)
    assert list(c.select_batch(3)) == [1, 0, 0]


def test_search_space_encoding():
    doc = {
        "hyperparameters": [
            {"name": "x0", "type": "uniform_float", "log": False, "lower": -512.0, "upper": 512.0, "default": -3.0, "q": None},
            {"name": "x1", "type": "uniform_float", "log": False, "lower": -512.0, "upper": 512.0, "default": -4.0, "q": None},
        ],
        "conditions": [], "forbiddens": [], "python_module_version": "0.4.17", "json_format_version": 0.2,
    }
    cs = search_space_to_config_space(doc, seed=1)
    assert [f.name for f in cs.values()] == ["x0", "x1"] and cs["x1"].default_value == -4.0
    hydra = {"hyperparameters": {"g": {"type": "normal_float", "mu": 9.8, "sigma": 1.0}}}
    cs2 = search_space_to_config_space(hydra)
    assert cs2["g"].mu == 9.8
    cs3 = FeatureSpace(name="myspace", space={"uniform_integer": (1, 10), "uniform_float": (1.0, 10.0),
                                              "categorical": ["a", "b", "c"], "constant": 1337})
    assert search_space_to_config_space(cs3) is cs3 and len(cs3) == 4
    assert search_space_to_config_space({"hyperparameters": {}}) is not None
    with pytest.raises(ValueError):
        search_space_to_config_space(3)


def test_get_context_bounds():
    lo, hi = get_context_bounds(["a", "b"], {"a": (-1.0, 2.0, float), "b": (0, np.inf, float)})
    assert list(lo) == [-1.0, 0.0] and hi[0] == 2.0 and np.isinf(hi[1])
