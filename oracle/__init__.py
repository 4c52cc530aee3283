"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's DCCRN train-step path.  Nothing in the product
package may import this; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do, and only as the checker / the timed CPU arm.
"""
