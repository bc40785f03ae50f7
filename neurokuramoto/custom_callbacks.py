"""The logging callbacks are the reference's own (aDBS_RL/agents/custom_callbacks.py, out of
scope here): re-exported when the reference checkout and stable-baselines3 are importable."""
try:
    from aDBS_RL.agents.custom_callbacks import EvalCallback_, TensorboardCallback  # noqa: F401
except Exception as exc:  # noqa: BLE001
    raise ImportError("neurokuramoto.custom_callbacks needs the reference's aDBS_RL package and "
                      "stable-baselines3 on the path") from exc
