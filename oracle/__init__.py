"""CPU oracle of the Long-VITA hot path.  TEST INFRASTRUCTURE ONLY: may be imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product
package long-vita_b200/."""
