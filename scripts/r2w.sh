python scripts/sweep_only.py --config 2 --patterns 64
python scripts/sweep_only.py --config 2
python scripts/sweep_phases.py --config 2
