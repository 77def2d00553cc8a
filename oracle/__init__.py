"""CPU oracle for the pair-force step. TEST INFRASTRUCTURE ONLY — never imported by hoomd_blue_b200."""
