"""Run a script with a Python-level watchdog: after N seconds dump every thread's stack and exit (finds WHERE a GPU job hangs).
Usage: python tools/run_with_watchdog.py SECONDS script.py [args...]"""
import faulthandler
import runpy
import sys

secs = int(sys.argv[1])
script = sys.argv[2]
sys.argv = sys.argv[2:]
faulthandler.dump_traceback_later(secs, exit=True)
runpy.run_path(script, run_name="__main__")
