"""CPU oracle package — test infrastructure only (see oracle/hk_oracle_lqng.c, oracle/hk_oracle_game.c)."""
