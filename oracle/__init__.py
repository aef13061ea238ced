"""CPU checkers for the self-collision path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package. The product (gpu-computing-course_b200/) never does.
"""
