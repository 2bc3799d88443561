"""CPU oracles for the triangulation path -- TEST INFRASTRUCTURE ONLY (see loop_oracle.py)."""
