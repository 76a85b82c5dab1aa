"""oracle/ — CPU restatements used ONLY as checkers (tests/, __graft_entry__.smoke(), and
bench.py's cpu_baseline / --impl reference legs). Nothing under hpc_b200/ imports this package."""
