"""CPU oracle (test infrastructure only -- see oracle/lewton_oracle.h)."""
