"""Community model hub (role of the reference's ``contrib/models/*``: ports built on the same base classes).  Each family here is a
thin definition over the engine's attention / MLP / decoder blocks and is checked against Hugging Face in tests/test_contrib_cpu.py."""
