"""CPU oracle of the vulkpy hot path: TEST INFRASTRUCTURE, never imported by vulkpy_b200."""
