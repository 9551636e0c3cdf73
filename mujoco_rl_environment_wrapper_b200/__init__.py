"""B200-native batched MuJoCoRL.step hot path (drop-in for MuJoCo_Gym.mujoco_rl.MuJoCoRL)."""
__version__ = "0.1.0"
