"""run_experiment.py (the reference's entry point, run_experiment.py:105-330) on the device path: `train` writes the
reference's artefacts, `eval` finds and rolls out the latest actor."""
import pickle

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env_name,obs_dim,act_dim", [("jvrc_walk", 37, 12), ("jvrc_step", 39, 12), ("h1", 35, 10)])
def test_train_then_eval(tmp_path, env_name, obs_dim, act_dim, capsys):
    import run_experiment as rx
    rx.main(["train", "--env", env_name, "--logdir", str(tmp_path), "--num-procs", "64", "--n-itr", "2", "--max-traj-len", "40",
             "--minibatch-size", "256", "--epochs", "1", "--seed", "3", "--eval-freq", "2"])
    runs = list(tmp_path.iterdir())
    assert len(runs) == 1 and runs[0].name.endswith("_" + env_name)
    names = sorted(p.name for p in runs[0].iterdir())
    # suffixed checkpoints at iteration 0 and every --eval-freq, the un-suffixed pair = the best evaluation so far
    assert [n for n in names if not n.startswith("events.out.tfevents")] == ["actor.pt", "actor_0.pt", "actor_1.pt", "critic.pt",
                                                                              "critic_0.pt", "critic_1.pt", "experiment.pkl"]
    ev = [n for n in names if n.startswith("events.out.tfevents")]
    assert len(ev) == 1                      # TensorBoard log with the reference's tags (rl/utils/logger.py:71-115)
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(str(runs[0]))
    acc.Reload()
    assert {"Loss/actor", "Loss/critic", "Loss/mirror", "Loss/imitation", "Train/mean_reward", "Train/mean_episode_length",
            "Train/mean_noise_std", "Time/fps", "Time/sample_time", "Time/optimize_time", "Time/total_elapsed",
            "Eval/mean_reward", "Eval/mean_episode_length"} <= set(
        acc.Tags()["scalars"])
    args = pickle.load(open(runs[0] / "experiment.pkl", "rb"))
    assert args.env == env_name and args.num_procs == 64
    actor = torch.load(runs[0] / "actor_1.pt", weights_only=False)
    assert actor(torch.zeros(3, obs_dim)).shape == (3, act_dim)        # a CPU copy (loadable on a box without a GPU)
    assert rx.get_latest_actor(runs[0]).name == "actor_1.pt"
    train_out = capsys.readouterr().out
    # the stdout lines scripts/benchmark_training.py:75-79 parses ARE the reference's metric definition
    import re
    for pat in (r"\*+ Iteration (\d+) \*+", r"Sampling took ([\d.]+)s for (\d+) steps", r"\|\s+Mean Eprew\s+\|\s+([\d.e+-]+)\s+\|",
                r"\|\s+Mean Eplen\s+\|\s+([\d.e+-]+)\s+\|", r"Total time elapsed: ([\d.]+)s.*fps=([\d.]+)",
                r"====EVALUATE EPISODE====\n\(Episode length:([\d.]+)\. Reward:([\d.e+-]+)\. Time taken:([\d.]+)s\)"):
        assert re.search(pat, train_out), pat
    assert re.search(r"Sampling took [\d.]+s for (\d+) steps", train_out).group(1) == str(64 * 40)
    episodes = rx.main(["eval", "--logdir", str(tmp_path), "--ep-len", "2"]) or []
    out = capsys.readouterr().out
    assert "episode(s) in 80 control steps" in out and "mean" in out


def test_unknown_env_and_unsupported_flags(tmp_path):
    import run_experiment as rx
    with pytest.raises(Exception, match="Check env name"):
        rx.main(["train", "--env", "cartpole", "--logdir", str(tmp_path), "--n-itr", "1"])
    with pytest.raises(NotImplementedError):
        rx.main(["train", "--env", "jvrc_walk", "--logdir", str(tmp_path), "--n-itr", "1", "--recurrent"])
