"""Oracle package: TEST INFRASTRUCTURE ONLY (see oracle/oracle.py)."""
