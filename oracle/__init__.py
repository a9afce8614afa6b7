"""CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/trtx_oracle.c header)."""
