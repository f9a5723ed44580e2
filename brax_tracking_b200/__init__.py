"""B200-native batched physics + tracking-reward step for the rodent / fruit-fly imitation envs."""
__version__ = "0.1.0"
