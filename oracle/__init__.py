"""CPU oracle package (test infrastructure only -- see oracle/tmc2_oracle.h)."""
