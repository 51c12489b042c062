#!/usr/bin/env bash
# Build the kernel library + an sdist/wheel and optionally copy it to a target (parity: reference deploy.sh).
# usage: ./deploy.sh [user@host:/path | /local/path]
set -euo pipefail
python -c "import __graft_entry__ as g; g.build()"
rm -rf dist build ./*.egg-info
python setup.py -q sdist bdist_wheel
ls -la dist
if [ "${1:-}" != "" ]; then
  case "$1" in
    *:*) scp dist/* "$1" ;;
    *)   mkdir -p "$1" && cp dist/* "$1" ;;
  esac
fi
