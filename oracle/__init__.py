"""CPU oracle for the CloudAAE hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under cloudaae_b200/ does.
"""
