"""TEST INFRASTRUCTURE -- CPU oracles for the hot path.  Never imported by the product
package; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it."""
