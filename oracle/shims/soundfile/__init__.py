"""Shim for `soundfile` (TEST INFRASTRUCTURE): imported by reference tests only."""
