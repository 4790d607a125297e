"""CPU oracle for the LGTEUN hot path — test infrastructure only (see lgteun_oracle.py header)."""
