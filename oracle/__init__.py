"""CPU oracle for the pointops hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import
this package.  The product (pytorch3d_pointops_b200) never does.
"""
