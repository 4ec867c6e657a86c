"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/scvx_oracle.cpp header)."""
