// TEST INFRASTRUCTURE ONLY (oracle/): empty stand-in, the reference header includes it without using it in the files built here.
#pragma once
