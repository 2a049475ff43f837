#!/usr/bin/env python3
"""Generate tests/golden/env_attributes.json by EXECUTING the reference's own method bodies:
`_setup_obs_normalization` of envs/jvrc/jvrc_walk.py, envs/jvrc/jvrc_step.py, envs/h1/h1_env.py and `_setup_mirror_indices` of
envs/jvrc/jvrc_base.py.  The modules import mujoco, so the functions are lifted out of the files by AST (as tools/gen_golden_h1.py
does for the noise functions) and run on a bare attribute holder that carries what they read (half_sitting_pose from the YAML
config, history_len, _get_num_external_obs)."""
import ast
import json
import os
import types

import numpy as np
import yaml

REF = os.environ.get("LHW_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def lift(rel, cls, fn):
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    node = next(f for c in tree.body if isinstance(c, ast.ClassDef) and c.name == cls for f in c.body
                if isinstance(f, ast.FunctionDef) and f.name == fn)
    node.returns = None
    mod = ast.Module(body=[node], type_ignores=[])
    ns = {"np": np}
    exec(compile(ast.fix_missing_locations(mod), rel, "exec"), ns)
    return ns[fn]


def main():
    out = {}
    jv = yaml.safe_load(open(os.path.join(REF, "envs/jvrc/configs/base.yaml")))
    h1 = yaml.safe_load(open(os.path.join(REF, "envs/h1/configs/base.yaml")))
    for name, rel, cls, half, n_ext in (("jvrc_walk", "envs/jvrc/jvrc_walk.py", "JvrcWalkEnv", jv["half_sitting_pose"], 8),
                                        ("jvrc_step", "envs/jvrc/jvrc_step.py", "JvrcStepEnv", jv["half_sitting_pose"], 10),
                                        ("h1", "envs/h1/h1_env.py", "H1Env", h1["half_sitting_pose"], 0)):
        holder = types.SimpleNamespace(half_sitting_pose=half, history_len=1, robot=types.SimpleNamespace(),
                                       _get_num_external_obs=lambda n=n_ext: n)
        lift(rel, cls, "_setup_obs_normalization")(holder)
        rec = dict(obs_mean=np.asarray(holder.obs_mean, dtype=float).tolist(), obs_std=np.asarray(holder.obs_std, dtype=float).tolist())
        if name != "h1":
            lift("envs/jvrc/jvrc_base.py", "JvrcBaseEnv", "_setup_mirror_indices")(holder)
            rec.update(mirrored_obs=[float(x) for x in holder.robot.mirrored_obs], mirrored_acts=[float(x) for x in holder.robot.mirrored_acts],
                       clock_inds=[int(x) for x in holder.robot.clock_inds])
        out[name] = rec
    json.dump(out, open(os.path.join(OUT, "env_attributes.json"), "w"))
    print("wrote env_attributes.json", {k: len(v["obs_mean"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
