"""Drop-in import path of the reference's `utils` package (reference main.py:12)."""
