"""CPU oracle for the all-pairs N-body hot path — TEST INFRASTRUCTURE ONLY (see reference_port.py)."""
