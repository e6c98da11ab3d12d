"""CPU oracle of the border hot path -- TEST INFRASTRUCTURE ONLY (see oracle/README.md)."""
