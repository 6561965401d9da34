"""CPU oracle for the evk kernels -- TEST INFRASTRUCTURE ONLY (see evk_oracle.c header)."""
