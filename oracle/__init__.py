"""CPU oracle of the hot path — TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; rtabmap_b200 never does.
"""
