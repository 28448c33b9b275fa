"""TEST INFRASTRUCTURE ONLY (see oracle/README.md) -- CPU restatement of the segment replay of input_data.cpp.

Follows the reference literally, globals and all, so that it can be read next to the source:
  IN:57-66    DISTANCE records, allLocation, pubLocation
  IN:78-122   subOdometryHandler: track point, travelled distance, next start location
  IN:244-446  main: two passes (long tracks, then short overlapping tracks), replay from the last start location,
              stop when the track is long enough, reset the SLAM nodes, queue of finished tracks, short-rest rule
ROS / rosbag are replaced by three callables:
  publish(bag_index, msg_index) -> (stamp, odometry or None)   odometry = (x, y, z, stamp) of /true_odometry_to_init
  control()                                                     IMControl{systemInited=false}
  slam_track(flag, [(x, y, z, t), ...])                         one IMTrack message
"""
import math

IMREST = 3.0  # IN:31


class _State:
    pass


def replay(messages_per_bag, long_distance, short_distance, overlap_distance, publish, control, slam_track, passes=(0, 1)):
    if not (long_distance > short_distance > overlap_distance > 0):  # IN:257
        raise ValueError("long > short > overlap > 0 required")
    g = _State()
    g.dSlamDistance = [long_distance, short_distance]  # IN:259-262
    g.dOverlapDistance = [0.0, overlap_distance]
    g.allLocation = []
    g.slamTrack = []
    g.slamTrackFlag = 0
    g.queue = []
    g.preOdometry = None
    g.dTotalDistance = 0.0
    g.published = g.lost = 0
    nbags = len(messages_per_bag)

    def new_loc(bag=0, msg=0, distance=0.0, timestamp=0.0):
        return {"bag": bag, "msg": msg, "distance": distance, "timestamp": timestamp}

    def publish_and_spin():
        stamp, odo = publish(g.nBagIndex, g.nMsgIndex - 1)
        g.published += 1
        if odo is None:
            return
        # ---- subOdometryHandler IN:78-122
        if odo[3] == stamp:
            g.slamTrack.append(tuple(odo))
            tmp = new_loc(g.nBagIndex, g.nMsgIndex)
            if g.preOdometry is None:
                tmp["distance"] = 0.0
            else:
                p = g.preOdometry
                tmp["distance"] = math.sqrt((odo[0] - p[0]) ** 2 + (odo[1] - p[1]) ** 2 + (odo[2] - p[2]) ** 2) + g.dTotalDistance
            tmp["timestamp"] = odo[3]
            g.preOdometry = odo
            if tmp["distance"] <= g.dSlamDistance[g.times] - g.dOverlapDistance[g.times]:
                g.pubLocation = tmp
            elif g.allLocation[-1]["timestamp"] != g.pubLocation["timestamp"]:
                g.allLocation.append(g.pubLocation)
            g.dTotalDistance = tmp["distance"]
        else:
            g.lost += 1

    for g.times in range(passes[0], passes[1] + 1):  # IN:272; passes = (first, last)
        g.pubLocation = new_loc()
        g.allLocation.append(g.pubLocation)
        g.nBagIndex = 0
        g.dTotalDistance = 0.0
        control()  # IN:281-285
        while g.nBagIndex < nbags:  # IN:287
            end = False
            g.nBagIndex = g.allLocation[-1]["bag"]
            while g.nBagIndex < nbags:  # IN:304
                g.nMsgIndex = 0
                for _ in range(messages_per_bag[g.nBagIndex]):
                    g.nMsgIndex += 1
                    if g.pubLocation["msg"] < g.nMsgIndex or g.pubLocation["bag"] < g.nBagIndex:  # IN:326
                        publish_and_spin()
                        if g.dTotalDistance > g.dSlamDistance[g.times]:  # IN:336
                            g.dTotalDistance = 0.0
                            end = True
                            break
                if end:
                    control()  # IN:350
                    break
                g.nBagIndex += 1
            g.slamTrackFlag = g.times  # IN:355-364
            g.queue.append((g.slamTrackFlag, list(g.slamTrack)))
            g.slamTrack = []
            g.preOdometry = None
            if len(g.queue) == 3:
                slam_track(*g.queue.pop(0))
        if len(g.allLocation) > 1 and g.dTotalDistance < g.dSlamDistance[g.times] / IMREST:  # IN:367
            tmp = g.allLocation[-2]
            g.queue = []
            g.slamTrack = []
            control()
            g.nBagIndex = tmp["bag"]
            while g.nBagIndex < nbags:
                g.nMsgIndex = 0
                for _ in range(messages_per_bag[g.nBagIndex]):
                    g.nMsgIndex += 1
                    if tmp["msg"] < g.nMsgIndex or tmp["bag"] < g.nBagIndex:
                        publish_and_spin()
                g.nBagIndex += 1
        if g.slamTrack:  # IN:428
            g.slamTrackFlag = g.times
            g.queue.append((g.slamTrackFlag, list(g.slamTrack)))
            g.slamTrack = []
        while g.queue:  # IN:435
            slam_track(*g.queue.pop(0))
        slam_track(g.slamTrackFlag, [])  # IN:441
        g.allLocation = []
    return {"published": g.published, "lost": g.lost}


class OracleSlam:
    """The CPU oracle's three nodes + transformMaintenance behind the publish/control callables."""

    def __init__(self, bags, stamps):
        import orc
        self.orc = orc
        self.bags, self.stamps = bags, stamps
        self.pipe = orc.Pipeline()
        self.tm = orc.TransformMaintenance()

    def control(self):
        self.pipe.reset()

    def publish(self, bag, msg):
        stamp = self.stamps[bag][msg]
        r = self.pipe.process(self.bags[bag][msg])
        if not r.odom_published:
            return stamp, None
        _, track = self.tm.odometry(list(r.odom), stamp)
        if r.mapping_ran:
            self.tm.aft_mapped(list(r.mapped), self.pipe.bef_mapped())
        return stamp, tuple(track)
