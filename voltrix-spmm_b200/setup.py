"""`pip install -e .` entry point, as in the reference (setup.py:6-11); the metadata lives in pyproject.toml."""
import setuptools

if __name__ == "__main__":
    setuptools.setup()
