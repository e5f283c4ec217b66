"""CPU oracle (test infrastructure only) -- see oracle/zoracle.c."""
