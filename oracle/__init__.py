"""CPU oracles (test infrastructure only -- see gp_oracle.py header)."""
