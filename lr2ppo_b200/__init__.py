"""lr2ppo_b200 — B200-native (sm_100a) implementation of the LR2PPO training / evaluation hot path.

Public surface mirrors the reference's module API:
  models:  Mlp, Actor, Critic, Reward, ActorCritic, Classifier, PairClassifier, XiT
  losses:  RankLoss, clipped_value_loss, ppo_policy_loss, pair_hinge_loss, smooth_l1_loss
  optim:   AdamW (FusedAdamW), get_linear_schedule_with_warmup, str2optimizer, str2scheduler
  ndcg:    AverageNDCGMeter
  ops:     one wrapper per C-ABI entry point of liblr2ppo_b200.so (include/lr2ppo_b200.h)
"""
__version__ = "0.1.0"
