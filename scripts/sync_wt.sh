#!/bin/bash
# Copy the build artefacts of the scratch worktree (.wt, where edits happen while a gpurun call is
# queued) into the main tree; sources travel through git (commit in .wt, merge here).
set -e
R=/root/repo
cp -a $R/.wt/pathfinder_b200/csrc/_build/*.o $R/.wt/pathfinder_b200/csrc/_build/*.log $R/pathfinder_b200/csrc/_build/
cp -a $R/.wt/pathfinder_b200/libpfb200.so $R/pathfinder_b200/
cp -a $R/.wt/oracle/_build/* $R/oracle/_build/
touch $R/pathfinder_b200/csrc/_build/*.o $R/pathfinder_b200/libpfb200.so $R/oracle/_build/*.so
make -C $R/pathfinder_b200/csrc -n | head -2
