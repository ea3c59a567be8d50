"""CPU oracles for the co-event counting path.  TEST INFRASTRUCTURE ONLY (see ref_restatement.py)."""
