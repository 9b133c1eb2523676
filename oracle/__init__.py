"""TEST INFRASTRUCTURE - CPU oracle for the raster / raster_pullback! hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
