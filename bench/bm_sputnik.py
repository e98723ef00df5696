"""Sputnik baseline (reference: bench/bm_sputnik.py): the same RoDe eval driver as bm_rode.py, which times both."""
import runpy
import os

runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bm_rode.py"), run_name="__main__")
