"""CPU oracle for the preprocessing chain -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product package (road-vision-system_b200/) never does.
"""
