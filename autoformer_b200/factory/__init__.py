"""Drop-in model classes: module name == class name, as the reference's train scripts expect
(``getattr(importlib.import_module(f"factory.{name}"), name)``, train.py:45-47)."""
