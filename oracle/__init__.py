"""CPU oracle for the log-likelihood hot path — TEST INFRASTRUCTURE, never imported by starfish_b200."""
