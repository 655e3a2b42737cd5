"""TEST INFRASTRUCTURE ONLY: CPU oracle for the GigaPose hot path (see oracle/port.py header).
Nothing under gigapose_b200/ or src/ may import this package."""
