for d in 1 2 4 8 16; do echo "density $d"; RT_GRID_DENSITY=$d python profiles/profile_render.py C3 8 2 counters 2>&1 | grep -v "^build" ; done
