"""Twins of the reference's ``make_data`` package that the conversion path's callers rely on."""
