"""CPU oracle for the fly() path — test infrastructure only (see ld_oracle.c header)."""
