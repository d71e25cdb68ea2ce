"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference algorithms on the LiDAR hot path, used exclusively as the
checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under futuredet_b200/ imports this package.
"""
