"""TEST INFRASTRUCTURE: CPU oracle (restatement of the reference search path). See awry_oracle.h."""
