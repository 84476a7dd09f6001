"""Drop-in mirror of the hot-path pieces of the reference's ``train/siamese_regions.py``."""
