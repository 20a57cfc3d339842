"""TEST INFRASTRUCTURE ONLY — CPU oracle for the LR2PPO hot path (see oracle/README.md)."""
