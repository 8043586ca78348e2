"""TEST INFRASTRUCTURE ONLY: CPU oracles for the fiesta_b200 parity tests (see oracle/pyoracle.py)."""
