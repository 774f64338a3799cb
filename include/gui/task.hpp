// gui/task.hpp — Task (gui/task.hpp:57-105), the base class of both stereo classes.  Qt-free:
// the signals started/finished/progressUpdate/stageUpdate become std::function callbacks, and
// the cancel flag is atomic (the reference's plain bool is a data race).
#ifndef SR_GUI_TASK_HPP
#define SR_GUI_TASK_HPP
#include "util/precompiled.hpp"
#include <atomic>
class Task {
public:
    Task() : cancelled(false) {}
    virtual ~Task() {}
    virtual std::string title() const = 0;
    virtual int numSteps() const = 0;
    bool isCancelled() const { return cancelled.load(); }
    void cancel() { cancelled.store(true); }
    void run() {  // gui/task.cpp:27-33
        if (onStarted) onStarted();
        runTask();
        if (onFinished) onFinished();
    }
    std::function<void()> onStarted, onFinished;
    std::function<void(int)> onProgressUpdate;
    std::function<void(const std::string &)> onStageUpdate;
protected:
    virtual void runTask() = 0;
    void progressUpdate(int v) { if (onProgressUpdate) onProgressUpdate(v); }
    void stageUpdate(const std::string &s) { if (onStageUpdate) onStageUpdate(s); }
private:
    std::atomic<bool> cancelled;
};
#endif
