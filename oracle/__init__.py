"""CPU oracle (test infrastructure). Import only from tests/, __graft_entry__.smoke() and bench.py."""
