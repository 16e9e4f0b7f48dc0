"""CPU oracle for the MultiGrid/PLR hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (dcd_isaac_b200) never does.
"""
