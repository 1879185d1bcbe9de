"""orb_slam3_b200: B200-native ORB front-end, Hamming/projection matchers and
local-BA LM engine behind ORB-SLAM3's own interfaces."""
