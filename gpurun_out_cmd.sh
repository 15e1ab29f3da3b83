python -c "import __graft_entry__ as g; g.smoke(); import master_doko_reinforcement_learning_b200 as p; print(p.load_library().dk_version())"
