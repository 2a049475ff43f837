"""Actor / critic modules (rl/policies/ of the reference) and the checkpoint interchange with a reference checkout.

The reference saves and loads WHOLE pickled modules (`torch.save(module)`: rl/utils/checkpointer.py:36-52, loaded by
run_experiment.py:274-277 and rl/algos/ppo.py:69-82), so a checkpoint names the class by import path
(`rl.policies.actor.Gaussian_FF_Actor`, `rl.policies.critic.FF_V`) and carries the instance dictionary.  For the two
directions to work
  * the classes here use the reference's attribute / sub-module names, so state-dict keys and instance dictionaries match;
  * they carry the reference's import path in `__module__`, and `install_reference_aliases()` registers
    `rl.policies.{base,actor,critic}` as aliases of this package WHEN no real `rl` package is importable — then
    `torch.save` here writes files a reference checkout unpickles with its own classes, and `torch.load` here resolves
    the names stored in reference-trained files;
  * when a reference checkout IS on the path, `export_module` builds the reference's own class instead.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
import types

import torch

from . import actor, base, critic
from .actor import Actor, Gaussian_FF_Actor
from .base import Net, normc_fn
from .critic import FF_V, Critic

_REF = {"rl.policies.base": base, "rl.policies.actor": actor, "rl.policies.critic": critic}
for _name, _mod in _REF.items():
    for _obj in vars(_mod).values():
        if isinstance(_obj, type) and _obj.__module__ == _mod.__name__:
            _obj.__module__ = _name          # pickled as the reference's import path


def install_reference_aliases() -> bool:
    """Make `rl.policies.actor` / `.critic` / `.base` importable as aliases of this package unless a real `rl` package
    (a reference checkout on sys.path) already is.  Returns True when the aliases are (already) in place."""
    if isinstance(sys.modules.get("rl.policies.actor"), types.ModuleType):
        return getattr(sys.modules["rl.policies.actor"], "Gaussian_FF_Actor", None) is Gaussian_FF_Actor
    try:
        real = importlib.util.find_spec("rl") is not None
    except (ImportError, ValueError):
        real = False
    if real:
        return False
    pkg = types.ModuleType("rl")
    pkg.__path__ = []          # a package, so that `import rl.policies.actor` walks through sys.modules
    pol = types.ModuleType("rl.policies")
    pol.__path__ = []
    pkg.policies = pol
    pol.base, pol.actor, pol.critic = base, actor, critic
    pol.Gaussian_FF_Actor, pol.FF_V = Gaussian_FF_Actor, FF_V
    sys.modules.update({"rl": pkg, "rl.policies": pol, **_REF})
    return True


def export_module(module: torch.nn.Module) -> torch.nn.Module:
    """A self-contained CPU copy of `module` for `torch.save`: built from whatever class `rl.policies.*` resolves to (this
    package through the alias, or the reference's own class when its checkout is on the path), with freshly allocated
    parameters — the training copy's parameters are views of one flat buffer, and pickling a view would store the whole
    buffer (actor + critic) in every file."""
    install_reference_aliases()
    ref_mod = "rl.policies.actor" if isinstance(module, Gaussian_FF_Actor) else "rl.policies.critic"
    cls = getattr(importlib.import_module(ref_mod), type(module).__name__)
    layers = tuple(lin.out_features for lin in (module.actor_layers if isinstance(module, Gaussian_FF_Actor) else module.critic_layers))
    if isinstance(module, Gaussian_FF_Actor):
        out = cls(module.state_dim, module.action_dim, layers=layers, init_std=0.2, learn_std=module.learn_std, bounded=module.bounded)
    else:
        out = cls(module.critic_layers[0].in_features, layers=layers)
    out.load_state_dict({k: v.detach().to("cpu", copy=True) for k, v in module.state_dict().items()})
    for name in ("obs_mean", "obs_std", "stds"):
        v = getattr(module, name, None)
        if torch.is_tensor(v) and not isinstance(v, torch.nn.Parameter):
            setattr(out, name, v.detach().to("cpu", copy=True))
    out.train(module.training)
    return out


install_reference_aliases()
