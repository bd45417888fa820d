"""CPU oracle for the LAPS hot path — test infrastructure only (see laps_oracle.py header)."""
