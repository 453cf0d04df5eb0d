#pragma once
#define GIT_BRANCH "oracle"
#define GIT_COMMIT_HASH "b458f57"
