"""CPU oracle (test infrastructure only). See oracle/oicr_plus_ref.py."""
