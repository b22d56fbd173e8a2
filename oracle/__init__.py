"""TEST INFRASTRUCTURE — CPU oracle of the reference's pileup hot path. Not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
