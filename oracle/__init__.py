"""CPU oracle (test infrastructure only) -- see mica_oracle.py."""
