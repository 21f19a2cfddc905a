"""Compatibility shims a Neuromah maintainer can drop into the reference tree."""
