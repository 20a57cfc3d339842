"""Stand-in for h5py on images that do not ship it: the read subset the LR2PPO scripts use, served from
`<file>.d/` directories of .npy arrays (lr2ppo_b200/h5shim.py)."""
from lr2ppo_b200.h5shim import Dataset, File, Group  # noqa: F401
