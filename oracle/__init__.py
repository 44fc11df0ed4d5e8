"""CPU oracle for the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; nothing under vidar_b200/ does (tests/test_layout.py checks).
"""
