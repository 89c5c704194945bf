"""Extract the golden vectors the reference holds for the context path from its example
notebooks (the only numeric outputs in the reference tree) into
tests/golden/notebook_goldens.json.

Run in the build container (needs /root/reference):  python tools/extract_notebook_goldens.py
"""
import ast
import json
import os
import re
import sys

REF = os.environ.get("CARL_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "notebook_goldens.json")


def cell_text(nb, idx):
    c = nb["cells"][idx]
    out = ""
    for o in c.get("outputs", []):
        t = o.get("text") or o.get("data", {}).get("text/plain") or ""
        out += "".join(t)
    return "".join(c["source"]), out


def first_dict(text):
    line = next(ln for ln in text.splitlines() if ln.startswith("{0:"))
    return ast.literal_eval(line)


def main():
    g = {}
    nb = json.load(open(os.path.join(REF, "examples/sample_contexts_with_brax.ipynb")))
    src, out = cell_text(nb, 5)
    assert 'NormalFloatContextFeature("gravity", mu=9.8, sigma=1' in src and "seed = 0" in src
    g["ant_gravity_normal_seed0_n5"] = {str(k): v for k, v in first_dict(out).items()}
    _, out = cell_text(nb, 3)
    g["ant_feature_names"] = ast.literal_eval(re.search(r"names for Ant: (\[.*\])", out).group(1))
    g["ant_default_context"] = ast.literal_eval(re.search(r"Default context for Ant: (\{.*\})", out).group(1))
    g["ant_friction_bounds"] = list(ast.literal_eval(re.search(r"friction in Ant: (\(.*\))", out).group(1)))
    ids = []
    for idx in (7, 9, 11):
        _, out = cell_text(nb, idx)
        ids.append(int(re.search(r"Current context ID: (\d+)", out).group(1)))
    g["round_robin_ids_reset_reset_set4"] = ids

    nb = json.load(open(os.path.join(REF, "examples/brax_with_goals.ipynb")))
    src, out = cell_text(nb, 1)
    assert "target_distance" in src and "target_direction" in src
    g["ant_target_seed0_n5"] = {str(k): v for k, v in first_dict(out).items()}
    src, out = cell_text(nb, 4)
    assert "goal_position_x" in src
    d = first_dict(out)
    g["pusher_goal_xy_seed0_n5"] = {
        str(k): {"goal_position_x": v["goal_position_x"], "goal_position_y": v["goal_position_y"]}
        for k, v in d.items()
    }
    _, out = cell_text(nb, 3)
    arr = re.search(r"Array\(\[(.*?)\], dtype=float32\)", out, re.S).group(1)
    g["ant_obs_after_one_random_step"] = [float(x) for x in arr.replace("\n", " ").split(",")]
    # cell 5: CARLBraxPusher (spring backend, default context) after reset() and ONE step with an unrecorded random
    # action: the printed observation f32[23] and reward -- the only output of a Brax body's step in the reference tree
    src, out = cell_text(nb, 5)
    assert "env.step(action)" in src
    arr = re.search(r"Array\(\[(.*?)\], dtype=float32\)", out, re.S).group(1)
    g["pusher_obs_after_one_random_step"] = [float(x) for x in arr.replace("\n", " ").split(",")]
    g["pusher_reward_after_one_random_step"] = float(re.search(r"\}\}\n(-?\d+\.\d+)", out).group(1))
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", os.path.abspath(OUT), "keys:", list(g))


if __name__ == "__main__":
    sys.exit(main())
