"""Drop-in mirror of the reference's ``test/`` helpers on the hot path."""
