"""CPU oracle of the TriFinger MDP hot path — TEST INFRASTRUCTURE (see trifinger_oracle.py).

Imported only by tests/, __graft_entry__.smoke() and the CPU legs of bench.py; never by the product package.
"""
