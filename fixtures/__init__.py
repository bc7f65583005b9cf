"""TEST/BENCH FIXTURES: synthetic texts and reference-format (.awry v1) index builders."""
