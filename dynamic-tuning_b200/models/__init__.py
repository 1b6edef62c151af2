"""Drop-in replacement of the reference's `models` package for the DyT hot path: put
`dynamic-tuning_b200/` ahead of the reference checkout on PYTHONPATH and `main_image.py`,
`main_vtab.py`, `speed.py` import these modules unchanged."""
