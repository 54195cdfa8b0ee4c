"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference's differentiable shading path, used ONLY as the checker by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
iris_b200/ imports it.  See oracle/README.md.
"""
