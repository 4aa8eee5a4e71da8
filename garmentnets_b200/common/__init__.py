"""Mirrors of the reference's ``common`` helpers that sit directly behind the hot path."""
