"""CPU oracle (test infrastructure only).  See oracle/ds_oracle.py."""
