"""CPU oracle (test infrastructure).  See gae_oracle.py for the scope rules."""
